#!/bin/bash
# full-size cross-engine check: fp32 sweep, three-product, two-product + refinement, single product + refinement
mkdir -p gpurun_out
for e in "sweep 0" "umma 3" "umma 2" "single 0"; do
  set -- $e
  timeout 600 python bench.py --engine $1 --nsplit $2 --steps 1 --warmup 3 --no-train --no-cpu --no-extras > gpurun_out/r02ad_$1_$2.json 2> gpurun_out/r02ad_$1_$2.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02ad_$1_$2.json'))
print('$1 $2', d['detail']['engine'], 'ms', round(d['ms_per_step'],1), 'cands', d['detail']['band_candidates_last_step'], d['rank_checksum'])
PY
done
