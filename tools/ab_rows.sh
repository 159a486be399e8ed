#!/bin/bash
# A/B of library builds on the row-block kernel: tools/ab_rows.sh lib1.so lib2.so ...
for lib in "$@"; do for w in slab c3 c5; do SQ_DEBUG=1 SQ_ROWS=1 SQ_LIBRARY=$PWD/stochquant_b200/$lib timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err; python -c "
import json;d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]);print('$lib $w',round(d['value']/1e9,1),'kernel us',round(d['roofline']['avg_launch_us'],1),d['clocks']['reasons'])"; grep "rows:" gpurun_out/ab.err | head -1; done; done
