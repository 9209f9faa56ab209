#!/bin/bash
# launch list of config 2 (HolE d = 150, pairwise, AdaGrad) training minibatches
mkdir -p gpurun_out
SKGE_EPOCHS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/r02by_cfg2_launches.csv python profiles/exp_configs.py cfg2 > gpurun_out/r02by_cfg2.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r02by_cfg2.log
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02by_cfg2_launches.csv')) if len(r) > 5]
h = rows[0]; ik = h.index('Kernel Name'); iv = h.index('Metric Value'); iu = h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    t = float(r[iv].replace(',', '')); t = t / 1e3 if r[iu] in ('ns', 'nsecond') else t
    a = agg.setdefault(r[ik][:70], [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
print('total us', tot, 'launches', len(rows) - 1)
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:30]: print('%-72s %4d %9.1f us %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
PY
