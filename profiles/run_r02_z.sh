#!/bin/bash
# the reference arm exactly as the driver launches it (N = 1 default flags, N = 2 under torchrun)
mkdir -p gpurun_out
( time python bench.py --impl reference ) > gpurun_out/r02z_ref1.json 2> gpurun_out/r02z_ref1.err; echo "rc=$?"
tail -1 gpurun_out/r02z_ref1.json | cut -c1-400; grep real gpurun_out/r02z_ref1.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/r02z_ref2.json 2> gpurun_out/r02z_ref2.err; echo "rc=$?"
grep -c impl gpurun_out/r02z_ref2.json; tail -1 gpurun_out/r02z_ref2.json | cut -c1-200; grep real gpurun_out/r02z_ref2.err
