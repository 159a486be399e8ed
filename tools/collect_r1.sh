#!/bin/bash
# one-GPU collection of the round's evidence: tests, bench lines, ncu launch list + full captures
TAG=${1:-r1}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/b.err || tail -3 gpurun_out/b.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/b.err
for w in c3 c5 slab c4s c4; do timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2>> gpurun_out/b.err; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_$TAG.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f.split("/")[-1], d.get("impl","ours"), round(d["value"]/1e9,3), "G/s ms/step", round(d["ms_per_step"],2), "kernel us", round(r.get("avg_launch_us",0)), "frac", round(r.get("frac",0),3), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks",{}).get("reasons"))
    except Exception as e: print(f, "ERR", e)
PY
# ncu: launch list of the default bench command, then full captures of the dominant kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resident2d -s 3 -c 1 -o gpurun_out/prof_c2_resident_$TAG python bench.py --workload c2 --steps 1 --warmup 3 --loops 100 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lattice_march -s 41 -c 1 -o gpurun_out/prof_c3_march_$TAG python bench.py --workload c3 --steps 1 --warmup 3 --loops 20 --no-cpu-baseline --no-e2e > gpurun_out/ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lattice_march|find_events" -s 12 -c 3 -o gpurun_out/prof_c4s_ring_$TAG python bench.py --workload c4s --steps 1 --warmup 3 --loops 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu4.log 2>&1
for f in gpurun_out/ncu2.log gpurun_out/ncu3.log gpurun_out/ncu4.log; do tail -n 1 $f; done
