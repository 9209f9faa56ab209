#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --engine single --steps 2 --warmup 3 --no-train --no-cpu --no-extras --test-triples 12800"
ncu --set full --clock-control none --import-source on -k regex:rank_single -s 3 -c 1 -o gpurun_out/r02u_single $CMD > gpurun_out/r02u_ncu.log 2>&1
echo "ncu rc=$?"
