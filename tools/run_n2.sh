#!/bin/bash
# N=2: default bench line (c2 replicas + ring block), the multi-process ring test, NVLink counters
out=gpurun_out/r02; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_default_n2.json 2> $out/bench_default_n2.err; echo rc=$?; tail -c 1500 $out/bench_default_n2.json; echo
timeout 300 python -m pytest tests -m gpu -q -k "multi_process" > $out/gputests_n2_ring.log 2>&1; tail -3 $out/gputests_n2_ring.log
SQ_ROWS=1 timeout 300 python -m pytest tests/test_gpu_slab.py -m gpu -q > $out/gputests_n2_rows.log 2>&1; tail -2 $out/gputests_n2_rows.log
