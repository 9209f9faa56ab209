#!/bin/bash
# round 2, first GPU call: parity of the new refine kernel (both CTA groupings) + a short bench of each
mkdir -p gpurun_out
export PYTHONPATH=.
for cg in 1 2; do
  echo "=== tests cg=$cg"
  timeout 600 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -k "refine or tensor_core" --deselect "tests/test_gpu_ranking.py::test_refine_engine_on_a_million_entities" -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r02a_tests_all.log
  echo "=== bench cg=$cg"
  SKGE_RANK_CG=$cg timeout 300 python bench.py --steps 5 --warmup 3 --no-train --no-cpu > gpurun_out/r02a_bench_cg$cg.json 2> gpurun_out/r02a_bench_cg$cg.err
  echo "rc=$?"; cat gpurun_out/r02a_bench_cg$cg.json | cut -c1-1500; tail -5 gpurun_out/r02a_bench_cg$cg.err

done
echo "=== 1M-entity test"
timeout 600 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -k "million" -p no:cacheprovider 2>&1 | tail -8
