#!/bin/bash
# register-resident d = 256 transforms in the spectral update: parity tests, then the config-5-size step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 300 python profiles/exp_train.py hole 4 2>&1 | tail -5
