#!/bin/bash
mkdir -p gpurun_out
for ns in 0 2; do
for wl in cfg2; do
  timeout 300 python bench.py --workload $wl --nsplit $ns --no-train --no-cpu --no-extras --steps 10 > gpurun_out/r02w_${wl}_$ns.json 2> gpurun_out/r02w_${wl}_$ns.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02w_${wl}_$ns.json'))
print('$wl nsplit $ns value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],d['detail']['engine'],d['rank_checksum'])
PY
done; done
