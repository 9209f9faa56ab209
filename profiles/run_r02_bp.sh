#!/bin/bash
# compact sweep epilogue (predicate masks + one push loop per tile): ranking parity suite, then configs 1 and 4
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ranking.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
for w in cfg1 cfg4; do
timeout 300 python bench.py --workload $w --no-train --no-cpu --no-extras --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', 'ms/pass', d['ms_per_step'], 'launch ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], d['rank_checksum'])"
done
