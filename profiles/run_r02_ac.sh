#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rank_rescore -s 4 -c 2 python bench.py --no-train --no-cpu --no-extras --steps 1 2>&1 | grep -E "gpu__time" | head -4
timeout 300 python bench.py --no-train --no-cpu --no-extras --steps 5 > gpurun_out/r02ac_cfg5.json 2> gpurun_out/r02ac_cfg5.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02ac_cfg5.json'))
print('cfg5 value',d['value'],'ms',d['ms_per_step'],'launch_ms',d['roofline']['launch_ms'],'other_ms',d['ms_per_step']-d['roofline']['launch_ms'],d['rank_checksum']['sum_filtered'])
PY
