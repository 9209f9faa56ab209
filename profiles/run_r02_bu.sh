#!/bin/bash
# checkpoint: full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02bu_bench.json 2> gpurun_out/r02bu_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02bu_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['rank_checksum'])
t=d.get('train',{})
print('train',t.get('value'),t.get('roofline',{}).get('frac'),t.get('roofline',{}).get('minibatch_ms'),t.get('epoch_times_s'), t.get('error'))
print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d.get('extra',{}).get('training',{}).items()})
print({k:(v.get('value'), v.get('roofline',{}).get('frac')) for k,v in d.get('extra',{}).get('ranking',{}).items()})
print(d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('gpu_ranks_equal_reference'))
PY
