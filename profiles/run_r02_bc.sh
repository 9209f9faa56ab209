#!/bin/bash
# bulk-TMA staged segmented update: parity, then A/B timings (HolE and TransE config-5-size steps)
mkdir -p gpurun_out
L=scikit-kge_b200/lib
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
for v in old bulk83 bulk64; do
  cp $L/variants/$v.so $L/libskge_b200.so
  echo "== $v"
  timeout 300 python profiles/exp_train.py hole 4 2>&1 | tail -2
  timeout 300 python profiles/exp_train.py transe 4 2>&1 | tail -2
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"seg_reduce" -s 1 -c 1 --csv --log-file gpurun_out/r02bc_$v.csv python profiles/exp_train.py hole 2 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02bc_$v.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]
for r in rows[hi+1:]:
    d=dict(zip(h,r))
    print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'])
PY
done
cp $L/variants/bulk83.so $L/libskge_b200.so
