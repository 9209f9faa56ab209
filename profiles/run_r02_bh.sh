#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "large_minibatch or twin_rows or frequency_domain or bit_reproducible or wn18_shaped or hot_rows or pairwise" 2>&1 | tail -15
timeout 300 python profiles/exp_train.py hole 4 2>&1 | tail -3
VARIANTS=${V:-shared} FINAL=${V:-shared} bash profiles/run_r02_be.sh
