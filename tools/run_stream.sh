#!/bin/bash
# streaming-kernel check: parity tests on the 3-D / 4-D lattices, the ring tests, benches of c3 / slab / c5.
# usage: run_stream.sh tag [ncu]  (env passes through: SQ_LIBRARY, SQ_NO_TILE ...)
tag=$1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_timed_configs.py tests/test_gpu_slab.py -x -q -k "not resident and not c2" > gpurun_out/${tag}_tests.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
for w in c3 slab c5; do
timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$w.json").read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print("${tag} $w: value %.1f G/s  e2e %.1f  frac %.3f  kernel_us %.1f (%s)  events %s clocks %s" % (d["value"]/1e9, (e.get("value") or 0)/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["kernel"], d.get("rng_events_replayed"), d["clocks"]["sm_mhz"]))
except Exception as e:
    print("bench $w failed", e); print(open("gpurun_out/${tag}_bench_$w.err").read()[-1500:])
PY
done
if [ -n "$2" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lattice_tile --launch-skip 10 -c 1 -o gpurun_out/${tag}_ncu -f python bench.py --workload slab --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
fi
