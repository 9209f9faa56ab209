#!/bin/bash
# relation-grouped RESCAL logistic kernel + in-place minibatch assembly: parity, then config 3 through fit()
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "rescal or logistic" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_trainer.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 300 python profiles/exp_configs.py cfg3 2>&1 | tail -1
