#!/bin/bash
mkdir -p gpurun_out
python profiles/exp_train.py hole 4 2>&1 | tail -6
python profiles/exp_train.py transe 4 2>&1 | tail -6
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02k_train_hole.csv python profiles/exp_train.py hole 2 > gpurun_out/r02k_ncu.log 2>&1
echo "ncu rc=$?"
# also: sweep retest (kend) and the full ranking test file
timeout 900 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
for wl in cfg4 cfg1; do
  timeout 300 python bench.py --workload $wl --no-train --no-cpu --no-extras --steps 5 > gpurun_out/r02k_${wl}.json 2> gpurun_out/r02k_${wl}.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02k_${wl}.json'))
print('$wl value',d['value'],'ms',d['ms_per_step'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],d['rank_checksum'])
PY
done
