#!/bin/bash
# ncu --set full (with source) of the staged pair kernel and the bulk update kernel of one config-5-size HolE step
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"seg_reduce_bulk|hole_pair_spec4" -s 2 -c 2 -f -o gpurun_out/r02bi_train python profiles/exp_train.py ${MODEL:-hole} 2 > gpurun_out/r02bi.log 2>&1
tail -3 gpurun_out/r02bi.log
ls -la gpurun_out/
