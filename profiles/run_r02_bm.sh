#!/bin/bash
# ncu --set full of every kernel of one config-5-size HolE minibatch step (own kernels and the CUB sorts / scan)
mkdir -p gpurun_out
timeout 300 python profiles/exp_train.py hole 3 2>&1 | tail -2
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"seg_|hole_pair|build_keys|scatter_heads|iota|sample_corrupt|RadixSort|DeviceScan" -s 79 -c 20 -f -o gpurun_out/r02bm_train_hole python profiles/exp_train.py hole 3 > gpurun_out/r02bm_ncu.log 2>&1
echo "ncu rc=$?"
tail -2 gpurun_out/r02bm_ncu.log
ncu -i gpurun_out/r02bm_train_hole.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
k=h.index('Kernel Name'); t=h.index('gpu__time_duration.sum')
for row in r[2:]: print(row[k][:60], row[t])
"
