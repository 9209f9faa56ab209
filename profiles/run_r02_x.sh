#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ranking.py tests/test_gpu_trainer.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02x_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for wl in cfg5 cfg2; do
  timeout 300 python bench.py --workload $wl --no-train --no-cpu --no-extras --steps 5 > gpurun_out/r02x_${wl}.json 2> gpurun_out/r02x_${wl}.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02x_${wl}.json'))
print('$wl value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launch_ms',d['roofline']['launch_ms'],'share',d['roofline']['kernel_share_of_step'],d['detail']['engine'],d['rank_checksum'])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:make_queries -c 6 python bench.py --no-train --no-cpu --no-extras --steps 1 2>&1 | grep -E "make_queries|gpu__time" | head -8
