#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --engine single --steps 2 --warmup 3 --no-train --no-cpu --no-extras --test-triples 12800"
$CMD > gpurun_out/r02j_plain.json 2> gpurun_out/r02j_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:rank_single -s 3 -c 1 -o gpurun_out/r02j_single_cg2 $CMD > gpurun_out/r02j_ncu.log 2>&1
echo "ncu rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02j_plain.json'))
print('value',d['value'],'ms',d['ms_per_step'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'])
PY
