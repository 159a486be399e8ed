#!/bin/bash
# A/B of tuning knobs on one workload: run_stream4.sh tag workload "ENV=.. ENV=.." ...
tag=$1; w=$2; shift 2
for cfg in "$@"; do
env $cfg timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$w.json").read().strip().splitlines()[-1])
    print("$cfg $w: value %.1f G/s  frac %.3f  kernel_us %.1f clocks %s" % (d["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("bench $w failed", e); print(open("gpurun_out/${tag}_$w.err").read()[-1500:])
PY
grep -E "ptile|rows:" gpurun_out/${tag}_$w.err | head -1
done
