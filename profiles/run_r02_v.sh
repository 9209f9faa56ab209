#!/bin/bash
mkdir -p gpurun_out
L=scikit-kge_b200/lib
cp $L/libskge_b200.so $L/orig.so
for v in orig alt_sleep512 alt_sleep2048; do
  cp $L/$v.so $L/libskge_b200.so
  for eng in umma single; do
  timeout 300 python bench.py --engine $eng --nsplit $([ $eng = umma ] && echo 2 || echo 0) --steps 5 --warmup 3 --no-train --no-cpu --no-extras > gpurun_out/r02v_$v_$eng.json 2> gpurun_out/r02v_$v_$eng.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02v_$v_$eng.json'))
print('$v $eng value',d['value'],'ms',d['ms_per_step'],'clk',d['clocks']['sm_mhz'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],d['rank_checksum']['sum_filtered'])
PY
  done
done
cp $L/orig.so $L/libskge_b200.so
