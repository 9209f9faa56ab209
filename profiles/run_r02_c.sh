#!/bin/bash
mkdir -p gpurun_out
for cg in 2 1; do
  echo "=== tests cg=$cg"
  SKGE_TEST_CG=$cg timeout 900 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -k "refine or tensor_core" -p no:cacheprovider 2>&1 | tail -5
  echo "=== bench cg=$cg"
  SKGE_RANK_CG=$cg timeout 300 python bench.py --steps 5 --warmup 3 --no-train --no-cpu > gpurun_out/r02c_bench_cg$cg.json 2> gpurun_out/r02c_bench_cg$cg.err
  echo "rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/r02c_bench_cg$cg.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],'cands',d['config']['band_candidates_last_step'])
PY
  tail -3 gpurun_out/r02c_bench_cg$cg.err
done
