#!/bin/bash
# final ncu --set full captures: L1 sweep (config 4) and single-product kernel (25.6 k queries of config 5)
mkdir -p gpurun_out
CMD="python bench.py --workload cfg4 --no-train --no-cpu --no-extras --steps 2 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:rank_sweep_tma -s 3 -c 1 -o gpurun_out/r02ag_sweep $CMD > gpurun_out/r02ag_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
CMD="python bench.py --engine single --steps 2 --warmup 3 --no-train --no-cpu --no-extras --test-triples 12800"
ncu --set full --clock-control none --import-source on -k regex:rank_single -s 3 -c 1 -o gpurun_out/r02ag_single $CMD > gpurun_out/r02ag_ncu_single.log 2>&1; echo "ncu single rc=$?"
