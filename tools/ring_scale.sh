#!/bin/bash
# usage: tools/ring_scale.sh N tag   (inside a gpurun --gpus N call) -- ring test + ring/independent benches at N GPUs
N=$1; TAG=$2
timeout 300 python -m pytest tests/test_gpu_slab.py -x -q -k multi_process 2>&1 | tail -2
for w in c4 c4s c2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_n${N}_${TAG}.json 2> gpurun_out/b_${w}_n${N}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${w}_n${N}_${TAG}.json").read().strip().splitlines()[-1])
    print("$w n$N", round(d["value"]/1e9,1), "G/s  ms/step", round(d["ms_per_step"],2), "kernel us", round(d["roofline"]["avg_launch_us"]), "frac", round(d["roofline"]["frac"],3), d["config"].get("slab"), d["clocks"]["reasons"])
except Exception as e:
    print("$w n$N failed", e); print(open("gpurun_out/b_${w}_n${N}.err").read()[-800:])
PY
done
