#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-train --no-cpu --no-extras"
ncu --set full --clock-control none --import-source on -k regex:rank_rescore -s 4 -c 2 -o gpurun_out/r02aa_rescore $CMD > gpurun_out/r02aa_ncu.log 2>&1
echo "ncu rc=$?"
