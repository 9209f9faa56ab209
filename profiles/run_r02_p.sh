#!/bin/bash
mkdir -p gpurun_out
SKGE_TEST_MODES=single timeout 900 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -p no:cacheprovider -k "refine or mixed" > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02p_tests.log | cut -c1-300
for eng in single; do
  timeout 300 python bench.py --engine $eng --steps 5 --warmup 3 --no-train --no-cpu --no-extras > gpurun_out/r02p_bench_$eng.json 2> gpurun_out/r02p_bench_$eng.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02p_bench_$eng.json'))
print('$eng value',d['value'],'ms',d['ms_per_step'],'clk',d['clocks']['sm_mhz'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],'cands',d['detail']['band_candidates_last_step'],d['rank_checksum']['sum_filtered'])
PY
  tail -3 gpurun_out/r02p_bench_$eng.err
done
