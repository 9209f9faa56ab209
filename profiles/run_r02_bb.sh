#!/bin/bash
# A/B of the spectral update: old shared-memory FFT vs register FFT (4 / 3 CTAs per SM); per-kernel times under ncu
mkdir -p gpurun_out
L=scikit-kge_b200/lib
for v in old ctas4 ctas3; do
  cp $L/variants/$v.so $L/libskge_b200.so
  echo "== $v"
  timeout 300 python profiles/exp_train.py hole 4 2>&1 | tail -2
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"seg_reduce|hole_pair_spec|seg_long" -s 4 -c 4 --csv --log-file gpurun_out/r02bb_$v.csv python profiles/exp_train.py hole 2 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02bb_$v.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]
for r in rows[hi+1:]:
    d=dict(zip(h,r))
    print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'])
PY
done
cp $L/variants/ctas4.so $L/libskge_b200.so
