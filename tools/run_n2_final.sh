#!/bin/bash
# last N = 2 check of the round: the new grouping test, the IPC ring test, the driver's multi-GPU line
out=gpurun_out/r02h; mkdir -p $out
timeout 100 python -m pytest tests -m gpu -q -k "finalize_grouping or multi_process" > $out/gputests_n2.log 2>&1; tail -2 $out/gputests_n2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > $out/bench_default_n2.json 2> $out/bench_default_n2.err; echo rc=$?; tail -c 600 $out/bench_default_n2.json; echo
