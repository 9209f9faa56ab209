#!/bin/bash
# ncu --set full of the refine kernel on a reduced query set (25.6 k queries, full 1 M-entity table)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-train --no-cpu --test-triples 12800"
SKGE_RANK_CG=1 $CMD > gpurun_out/r02b_plain.json 2> gpurun_out/r02b_plain.err && \
SKGE_RANK_CG=1 ncu --set full --clock-control none --import-source on -k regex:rank_refine -s 3 -c 1 -o gpurun_out/r02b_refine_cg1 $CMD > gpurun_out/r02b_ncu.log 2>&1
echo "rc=$?"; cut -c1-600 gpurun_out/r02b_plain.json; tail -3 gpurun_out/r02b_ncu.log
