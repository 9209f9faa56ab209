#!/bin/bash
# new bulk-TMA sweep: parity (ranking tests use the sweep engine as the reference engine), A/B against the cp.async kernel,
# config-sized parity test, issue-rate microbenchmark
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/issue_rates profiles/exp_issue_rates.cu && /tmp/issue_rates | tee gpurun_out/r02g_issue_rates.txt
timeout 900 python -m pytest tests/test_gpu_ranking.py tests/test_gpu_config_parity.py -x -q -m gpu -s -p no:cacheprovider > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?"; grep -E "quantiles|violations per|trained|passed|failed" gpurun_out/r02g_tests.log | cut -c1-500
for st in tma cp.async; do
for wl in cfg4 cfg1; do
  SKGE_SWEEP_STAGING=$st timeout 300 python bench.py --workload $wl --no-train --no-cpu --no-extras --steps 5 > gpurun_out/r02g_${wl}_${st}.json 2> gpurun_out/r02g_${wl}_${st}.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02g_${wl}_${st}.json'))
print('$wl $st value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],d['rank_checksum'])
PY
done; done
