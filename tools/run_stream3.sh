#!/bin/bash
# benches only: c3 / slab / c5 for the library and knobs in the environment
tag=$1
for w in c3 slab c5; do
timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$w.json").read().strip().splitlines()[-1])
    print("${tag} $w: value %.1f G/s  frac %.3f  kernel_us %.1f  events %s clocks %s" % (d["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d.get("rng_events_replayed"), d["clocks"]["sm_mhz"]))
except Exception as e:
    print("bench $w failed", e); print(open("gpurun_out/${tag}_bench_$w.err").read()[-1500:])
PY
done
