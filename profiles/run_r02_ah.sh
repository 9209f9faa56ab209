#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
n=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02ah_n$n.json 2> gpurun_out/r02ah_n$n.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ah_n$n.json').read().strip().splitlines()[-1])
print('n=$n value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'share',d['roofline']['kernel_share_of_step'],'launch_ms',d['roofline']['launch_ms'],d['rank_checksum'])
PY
