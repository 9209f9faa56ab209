#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_gpu_config_parity.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python profiles/exp_train.py hole 4 2>&1 | tail -2
timeout 300 python profiles/exp_train.py transe 4 2>&1 | tail -1
VARIANTS=keys FINAL=keys bash profiles/run_r02_be.sh 2>&1 | grep -v "RadixSort"
