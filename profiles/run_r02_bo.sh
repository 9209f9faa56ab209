#!/bin/bash
# parity suites after the mixed-radix change (rest of the run cut short by -x in run_r02_bn.sh), then an
# ncu --set full capture of the L1 sweep at config 1 (d = 50: 0.50 of the issue bound vs 0.74 at d = 200)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_gpu_config_parity.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
CMD="python bench.py --workload cfg1 --no-train --no-cpu --no-extras --steps 2 --warmup 3"
timeout 300 $CMD 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1', d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rank_sweep_tma -s 3 -c 1 -f -o gpurun_out/r02bo_sweep_cfg1 $CMD > gpurun_out/r02bo_ncu.log 2>&1; echo "ncu rc=$?"
