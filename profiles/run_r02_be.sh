#!/bin/bash
# per-kernel times of one config-5-size HolE step (own kernels + CUB sorts)
mkdir -p gpurun_out
L=scikit-kge_b200/lib
for v in ${VARIANTS:-rel3}; do
  cp $L/variants/$v.so $L/libskge_b200.so
  echo "== $v"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none -k regex:"seg_|hole_pair|build_keys|mark_heads|scatter_heads|iota|sample_corrupt|RadixSort|DeviceScan" -c 120 --csv --log-file gpurun_out/r02be_$v.csv python profiles/exp_train.py ${MODEL:-hole} 2 > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02be_$v.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]
acc=collections.OrderedDict()
for r in rows[hi+1:]:
    d=dict(zip(h,r))
    acc.setdefault((int(d['ID']), d['Kernel Name'][:48]), {})[d['Metric Name']] = d['Metric Value']
items=list(acc.items())
# last step = after the last sample_corrupt kernel
last=[i for i,((_,k),m) in enumerate(items) if 'sample_corrupt' in k][-1]
tot=0
for (i,k),m in items[last:]:
    t=float(m.get('gpu__time_duration.sum','0').replace(',',''))/1e3
    tot+=t
    print('%4d %-48s %8.1f us  rd %7.1f MB wr %7.1f MB  inst %s regs %s' % (i,k,t,float(m['dram__bytes_read.sum'].replace(',',''))/1e6,float(m['dram__bytes_write.sum'].replace(',',''))/1e6,m['smsp__inst_executed.sum'],m['launch__registers_per_thread']))
print('total %.1f us' % tot)
PY
done
cp $L/variants/${FINAL:-rel3}.so $L/libskge_b200.so
