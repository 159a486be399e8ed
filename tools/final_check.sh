#!/bin/bash
out=gpurun_out/r02f; mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -q -x) > $out/gputests.log 2>&1; tail -3 $out/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo rc=$?; python -c "
import json;d=json.loads(open('$out/bench_default.json').read().strip().splitlines()[-1]);print('default',round(d['value']/1e9,1),'e2e',round(d['e2e']['value']/1e9,1),'frac',round(d['roofline']['frac'],3),'traffic',d['roofline']['traffic'],d['clocks']['reasons'],'launches',d['gpu_launches']); print({k:round(v.get('value',0)/1e9,1) for k,v in d['extras'].items() if 'value' in v})"
timeout 400 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo rc=$?; tail -c 300 $out/bench_reference.json
for w in c3 slab; do timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_$w.json 2>$out/bench_$w.err; python -c "
import json;d=json.loads(open('$out/bench_$w.json').read().strip().splitlines()[-1]);print('$w',round(d['value']/1e9,1),'frac',round(d['roofline']['frac'],3),'kus',round(d['roofline']['avg_launch_us'],1),'traffic',d['roofline']['traffic'])"; done
