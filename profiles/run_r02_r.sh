#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02r_tests.log | cut -c1-300
python profiles/exp_train.py hole 4 2>&1 | tail -3
