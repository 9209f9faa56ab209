#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "large_minibatch or twin_rows or frequency_domain or bit_reproducible or wn18_shaped" 2>&1 | tail -15
