#!/bin/bash
# inside `gpurun --gpus N`: the driver's multi-GPU line (c2 replicas + the 256^4 ring block) and, at N = 8, configs[4] (4096 chains)
N=$1; out=gpurun_out/r02; mkdir -p $out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > $out/$2.json 2> $out/$2.err; echo "$2 rc=$?"; tail -c 900 $out/$2.json; echo; }
run 29521 bench_default_n$N --steps 10 --warmup 3
if [ "$N" = 8 ]; then run 29522 bench_c5_n8 --workload c5 --steps 5 --warmup 3; fi
