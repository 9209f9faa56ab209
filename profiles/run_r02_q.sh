#!/bin/bash
# ncu --set full captures for the bench's roofline.traffic fields (full-size launches), shipped commit
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-train --no-cpu --no-extras"
$CMD > gpurun_out/r02q_plain.json 2> gpurun_out/r02q_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:rank_refine -s 3 -c 1 -o gpurun_out/r02q_refine_full $CMD > gpurun_out/r02q_ncu.log 2>&1
echo "ncu refine rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"hole_pair_spec|seg_reduce|seg_long|sample_corrupt|build_keys|mark_heads|scatter_heads" -s 18 -c 9 -o gpurun_out/r02q_train_hole python profiles/exp_train.py hole 3 > gpurun_out/r02q_ncu_train.log 2>&1
echo "ncu train rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r02q_ncu_list.log 2>&1
echo "ncu list rc=$?"
