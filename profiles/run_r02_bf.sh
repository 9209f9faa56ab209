#!/bin/bash
# TMA-staged spectral pair kernel: parity, timing, per-kernel breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_gpu_config_parity.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 300 python profiles/exp_train.py hole 4 2>&1 | tail -3
VARIANTS=staged3 FINAL=staged3 bash profiles/run_r02_be.sh
