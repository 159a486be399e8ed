#!/bin/bash
# inside `gpurun --gpus 2`: NVLink byte counters of the update kernel of rank 0 of a two-process ring (rank 1 runs free).
# Only single-pass metrics: a replayed pass would meet flags that are already raised.
out=gpurun_out/r02k; mkdir -p $out
name=nv$$
timeout 70 python tools/nvlink_rank.py $name 1 2 > $out/rank1.log 2>&1 &
timeout 70 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum --clock-control none -k regex:lattice_ -c 12 --csv \
    --log-file $out/nvlink_rank0.csv python tools/nvlink_rank.py $name 0 2 > $out/rank0.log 2>&1
echo rc0=$?
wait
tail -2 $out/rank0.log $out/rank1.log
grep -c lattice_ $out/nvlink_rank0.csv; grep -i "nvl" $out/nvlink_rank0.csv | head -8
ncu --query-metrics 2>/dev/null | grep -i "^nvl" | head -30 > $out/nvl_metrics.txt; wc -l $out/nvl_metrics.txt
