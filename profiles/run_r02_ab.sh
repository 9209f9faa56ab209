#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ranking.py tests/test_gpu_config_parity.py tests/test_gpu_trainer.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02ab_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --no-train --no-extras --steps 5 > gpurun_out/r02ab_cfg5.json 2> gpurun_out/r02ab_cfg5.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02ab_cfg5.json'))
print('cfg5 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launch_ms',d['roofline']['launch_ms'],'share',d['roofline']['kernel_share_of_step'],d['rank_checksum'], d['cpu_baseline'].get('gpu_ranks_equal_reference'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rank_rescore -s 4 -c 2 python bench.py --no-train --no-cpu --no-extras --steps 1 2>&1 | grep -E "gpu__time" | head -4
