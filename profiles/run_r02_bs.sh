#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "rescal or logistic" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_trainer.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python profiles/exp_configs.py cfg3 2>&1 | tail -1
bash profiles/run_r02_bq.sh 2>&1 | head -16
