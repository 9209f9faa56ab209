#!/bin/bash
mkdir -p gpurun_out
SKGE_EPOCHS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rescal_logistic_grouped -s 120 -c 1 -f -o gpurun_out/r02bt_rescal python profiles/exp_configs.py cfg3 > gpurun_out/r02bt.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r02bt.log
