#!/bin/bash
# configs[0]: the reference's default run (taumain.py:101-132) through the drop-in tauhost.o, stdout to /dev/null
cd $GRAFT_REPO_ROOT
for frames in 500 5000; do
  s=$(date +%s.%N); ./tauhost.o 200 0.02 0.002 $frames 3 1.0 2 1 0 1000 0 gpurun_out/c1_end_$frames.txt 40 > gpurun_out/c1_stdout_$frames.txt; rc=$?; e=$(date +%s.%N)
  python - <<PY
import sys
dt=$e-$s; frames=$frames
lines=open("gpurun_out/c1_stdout_$frames.txt").read().splitlines()
print(f"C1 frames={frames} rc=$rc wall {dt:.2f} s  -> {frames/dt:.0f} frames/s, <= {frames*1000*200/dt/1e6:.1f} M site-updates/s (rejected frames stop early); last dtau {lines[-1].split('|')[-2].strip()}")
PY
done
rm -f gpurun_out/c1_stdout_*.txt
