#!/bin/bash
mkdir -p gpurun_out
SKGE_TEST_MODES=single timeout 900 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -p no:cacheprovider -k "refine or mixed" > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02i_tests.log | cut -c1-300
for cg in 2 1; do
  SKGE_RANK_CG=$cg timeout 300 python bench.py --engine single --steps 5 --warmup 3 --no-train --no-cpu --no-extras > gpurun_out/r02i_bench_cg$cg.json 2> gpurun_out/r02i_bench_cg$cg.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02i_bench_cg$cg.json'))
print('single cg$cg value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],'cands',d['detail']['band_candidates_last_step'],d['rank_checksum'])
PY
  tail -3 gpurun_out/r02i_bench_cg$cg.err
done
for wl in cfg4 cfg1; do
  timeout 300 python bench.py --workload $wl --no-train --no-cpu --no-extras --steps 5 > gpurun_out/r02i_${wl}.json 2> gpurun_out/r02i_${wl}.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02i_${wl}.json'))
print('$wl value',d['value'],'ms',d['ms_per_step'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],d['rank_checksum'])
PY
done
