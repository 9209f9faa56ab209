#!/bin/bash
# relation rows pre-reduced in the spectral pair kernel: parity, A/B timings, per-kernel times of one step
mkdir -p gpurun_out
L=scikit-kge_b200/lib
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_gpu_config_parity.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
for v in bulk64 rel3 rel2; do
  cp $L/variants/$v.so $L/libskge_b200.so
  echo "== $v"
  timeout 300 python profiles/exp_train.py hole 4 2>&1 | tail -2
done
for v in rel3; do
  cp $L/variants/$v.so $L/libskge_b200.so
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02bd_$v.csv python profiles/exp_train.py hole 2 > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02bd_$v.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]
acc=collections.OrderedDict()
for r in rows[hi+1:]:
    d=dict(zip(h,r))
    acc.setdefault((d['ID'], d['Kernel Name'][:60]), {})[d['Metric Name']] = d['Metric Value']
for (i,k),m in acc.items():
    print(i, k, m.get('gpu__time_duration.sum'), m.get('dram__bytes_read.sum'), m.get('dram__bytes_write.sum'), m.get('smsp__inst_executed.sum'))
PY
done
cp $L/variants/rel3.so $L/libskge_b200.so
