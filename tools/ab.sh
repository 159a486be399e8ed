#!/bin/bash
# A/B of library builds on the streaming workloads: tools/ab.sh lib1.so lib2.so ...
for lib in "$@"; do for w in slab c3 c5; do SQ_LIBRARY=$PWD/stochquant_b200/$lib timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err; python -c "
import json;d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]);print('$lib $w',round(d['value']/1e9,1),'kernel us',round(d['roofline']['avg_launch_us']),d['clocks']['reasons'])"; done; done
