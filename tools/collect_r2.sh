#!/bin/bash
# Round-2 evidence on ONE B200 (run through gpurun): GPU tests, the default bench line, the reference arm, the other
# workloads' lines, the ncu launch list of the default bench and full captures of the two dominant kernels.
out=gpurun_out/r02
mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -q) > $out/gputests.log 2>&1; tail -3 $out/gputests.log
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_c2_n1.json 2> $out/bench_c2_n1.err; tail -c 300 $out/bench_c2_n1.json; echo
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > $out/bench_c2_reference.json 2> $out/bench_c2_reference.err
for w in c3 c5 slab c2phi4; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > $out/bench_${w}_n1.json 2> $out/bench_${w}_n1.err; done
timeout 300 python bench.py --workload c4s --steps 5 --warmup 3 > $out/bench_c4s_ring1.json 2> $out/bench_c4s_ring1.err
timeout 300 python bench.py --workload c3 --math accurate --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_c3_accurate.json 2> $out/bench_c3_accurate.err
timeout 400 python bench.py --workload c1 > $out/bench_c1.json 2> $out/bench_c1.err; tail -c 400 $out/bench_c1.json; echo
# ncu: launch list of the default bench (its own command, short), then full captures
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $out/ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowres --launch-skip 3 -c 1 -o $out/ncu_c2_rowres -f python bench.py --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $out/ncu_c2.log 2>&1; tail -1 $out/ncu_c2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lattice_tile --launch-skip 30 -c 1 -o $out/ncu_c3_tile -f python bench.py --workload c3 --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $out/ncu_c3.log 2>&1; tail -1 $out/ncu_c3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lattice_tile --launch-skip 12 -c 1 -o $out/ncu_slab_tile -f python bench.py --workload slab --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $out/ncu_slab.log 2>&1; tail -1 $out/ncu_slab.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:find_events --launch-skip 4 -c 1 -o $out/ncu_finder -f python bench.py --workload c4s --steps 1 --warmup 3 --no-extras > $out/ncu_finder.log 2>&1; tail -1 $out/ncu_finder.log
SQ_ROWS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lattice_rows --launch-skip 12 -c 1 -o $out/ncu_slab_rows -f python bench.py --workload slab --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $out/ncu_slab_rows.log 2>&1; tail -1 $out/ncu_slab_rows.log
for w in c3 slab c5; do SQ_ROWS=1 timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_${w}_rowblock.json 2> $out/bench_${w}_rowblock.err; done
ls -la $out | head -60
