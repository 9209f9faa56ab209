#!/bin/bash
# mixed-radix (d / 2 = 2^a 3^b 5^c) spectral HolE step: parity suites, then d = 150 at config-5 size
# (spectral vs the direct O(d^2) kernel) and configs 1-4 through fit()
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_gpu_config_parity.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15
timeout 200 python profiles/exp_train.py hole 2 150 2>&1 | tail -2
SKGE_SPECTRAL=0 timeout 300 python profiles/exp_train.py hole 2 150 2>&1 | tail -2
timeout 200 python profiles/exp_train.py hole 2 200 2>&1 | tail -1
timeout 300 python profiles/exp_configs.py 2>&1 | tail -5
