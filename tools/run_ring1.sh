#!/bin/bash
# ring of one on the c4 quarter-size lattice (256^3 x 64): the finder + REBASE path on one GPU, against the plain context
tag=$1
timeout 300 python bench.py --workload c4s --steps 4 --warmup 3 --no-extras > gpurun_out/${tag}_c4s.json 2> gpurun_out/${tag}_c4s.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_c4s.json").read().strip().splitlines()[-1])
    print("${tag} c4s ring-of-1: value %.1f G/s frac %.3f kernel_us %.1f slab %s clocks %s" % (d["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["slab"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/${tag}_c4s.err").read()[-1500:])
PY
