#!/bin/bash
# A/B of the finalize grouping (sq_enqueue_step) inside one gpurun call, then the round's validation on the default
out=gpurun_out/r02g; mkdir -p $out
for w in small c3 slab c5; do
  for fb in 1 4 8; do SQ_FIN_BATCH=$fb timeout 120 python tools/exp_fin_batch.py $w; done
  if [ $w = c3 ]; then
    SQ_FIN_BATCH=4 SQ_PDL=0 timeout 120 python tools/exp_fin_batch.py $w
    SQ_FIN_BATCH=8 SQ_MARCH_R=4 timeout 120 python tools/exp_fin_batch.py $w   # 4096 smaller tiles: a shorter last wave
  fi
done 2>&1 | tee $out/fin_batch.log
(time timeout 600 python -m pytest tests -m gpu -q -x) > $out/gputests.log 2>&1; tail -3 $out/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 400 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo rc=$?; python -c "
import json;d=json.loads(open('$out/bench_default.json').read().strip().splitlines()[-1]);print('default',round(d['value']/1e9,1),'e2e',round(d['e2e']['value']/1e9,1),'frac',round(d['roofline']['frac'],3),'launches',d['gpu_launches']); print({k:round(v.get('value',0)/1e9,1) for k,v in d['extras'].items() if 'value' in v})"
