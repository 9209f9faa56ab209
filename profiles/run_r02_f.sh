#!/bin/bash
# full GPU test suite + the default bench line + the reference arm (short)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02f_tests.log
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02f_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks'])
print('roof',{k:d['roofline'][k] for k in ('frac','launch_ms','kernel_share_of_step','engine')})
print('checksum',d['rank_checksum'])
print('train',json.dumps(d.get('train'))[:600])
print('extra',json.dumps(d.get('extra'))[:3000])
print('cpu',json.dumps(d.get('cpu_baseline'))[:800])
PY
