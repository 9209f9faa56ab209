#!/bin/bash
# full GPU test suite + smoke + the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r02ae_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02ae_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02ae_bench.json 2> gpurun_out/r02ae_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02ae_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02ae_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks'])
print('roof',{k:d['roofline'][k] for k in ('frac','launch_ms','kernel_share_of_step','engine','traffic')})
print('checksum',d['rank_checksum'])
print('train',d['train']['value'],d['train']['roofline']['frac'],d['train']['roofline']['minibatch_ms'],d['train']['roofline']['traffic'])
for k,v in d['extra']['ranking'].items(): print(k, v.get('value'), v.get('ms_per_pass'), v.get('roofline',{}).get('frac'))
for k,v in d['extra']['training'].items(): print(k, v.get('value'), v.get('us_per_minibatch'))
print('cpu',json.dumps(d.get('cpu_baseline'))[:400])
PY
