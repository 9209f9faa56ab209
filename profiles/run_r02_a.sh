#!/bin/bash
# round 2, first GPU call: parity of the new refine kernel (both CTA groupings) + a short bench of each
mkdir -p gpurun_out
for cg in 1 2; do
  echo "=== tests cg=$cg"
  SKGE_TEST_CG=$cg timeout 900 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -k "refine or tensor_core" -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r02a_tests_cg$cg.log
  echo "=== bench cg=$cg"
  SKGE_RANK_CG=$cg timeout 300 python bench.py --steps 5 --warmup 3 --no-train --no-cpu > gpurun_out/r02a_bench_cg$cg.json 2> gpurun_out/r02a_bench_cg$cg.err
  echo "rc=$?"; cut -c1-1800 gpurun_out/r02a_bench_cg$cg.json; tail -5 gpurun_out/r02a_bench_cg$cg.err
done
