#!/bin/bash
# the shipped code under torchrun on 8 GPUs: rank checksum must equal the 1-GPU line's
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02bz_n8.json 2> gpurun_out/r02bz_n8.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02bz_n8.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['rank_checksum'])
PY
tail -c 300 gpurun_out/r02bz_n8.err
