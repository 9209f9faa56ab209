#!/bin/bash
mkdir -p gpurun_out
L=scikit-kge_b200/lib
cp $L/variants/blk64.so $L/libskge_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "large_minibatch or twin_rows or frequency_domain or bit_reproducible or wn18_shaped or hot_rows or pairwise" 2>&1 | tail -5
for v in shared blk64 blk83; do
  cp $L/variants/$v.so $L/libskge_b200.so
  echo "== $v"
  timeout 300 python profiles/exp_train.py hole 4 2>&1 | tail -2
  timeout 300 python profiles/exp_train.py transe 4 2>&1 | tail -1
done
VARIANTS="blk64 blk83" FINAL=blk64 bash profiles/run_r02_be.sh 2>&1 | grep "==\|seg_\|total"
