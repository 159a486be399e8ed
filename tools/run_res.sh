#!/bin/bash
# resident-kernel check: parity tests, bench, optional ncu capture.  usage: run_res.sh tag [ncu] ; env passes through
tag=$1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_timed_configs.py -x -q -k "resident or c2 or clamp or cold_start or free_field_ensemble" > gpurun_out/${tag}_tests.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("${tag}: value %.1f G/s  e2e %.1f  frac %.3f  kernel_us %.1f  ms/step %.3f events %s" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["ms_per_step"], d.get("rng_events_replayed")))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-2000:])
PY
if [ -n "$2" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowres --launch-skip 3 -c 1 -o gpurun_out/${tag}_ncu -f python bench.py --steps 1 --warmup 3 --loops 100 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
fi
