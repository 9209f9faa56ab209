#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload cfg4 --no-train --no-cpu --no-extras --steps 2 --warmup 3"
$CMD > gpurun_out/r02h_plain.json 2> gpurun_out/r02h_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:rank_sweep_tma -s 3 -c 1 -o gpurun_out/r02h_sweep_tma $CMD > gpurun_out/r02h_ncu.log 2>&1
echo "ncu rc=$?"
SKGE_SWEEP_STAGING=cp.async ncu --set full --clock-control none --import-source on -k regex:rank_sweep_kernel -s 3 -c 1 -o gpurun_out/r02h_sweep_old $CMD > gpurun_out/r02h_ncu_old.log 2>&1
echo "ncu old rc=$?"
