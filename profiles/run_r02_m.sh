#!/bin/bash
mkdir -p gpurun_out
SKGE_TEST_MODES=single timeout 900 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -p no:cacheprovider -k "refine or mixed" > gpurun_out/r02n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02n_tests.log | cut -c1-300
for cg in 2; do
  SKGE_RANK_CG=$cg timeout 300 python bench.py --engine single --steps 5 --warmup 3 --no-train --no-cpu --no-extras > gpurun_out/r02n_bench_cg$cg.json 2> gpurun_out/r02n_bench_cg$cg.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02n_bench_cg$cg.json'))
print('single cg$cg value',d['value'],'ms',d['ms_per_step'],'clk',d['clocks']['sm_mhz'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],'cands',d['detail']['band_candidates_last_step'],d['rank_checksum']['sum_filtered'])
PY
  tail -3 gpurun_out/r02n_bench_cg$cg.err
done
CMD="python bench.py --engine single --steps 2 --warmup 3 --no-train --no-cpu --no-extras --test-triples 12800"
ncu --set full --clock-control none --import-source on -k regex:rank_single -s 3 -c 1 -o gpurun_out/r02n_single_cg2 $CMD > gpurun_out/r02n_ncu.log 2>&1
echo "ncu rc=$?"
