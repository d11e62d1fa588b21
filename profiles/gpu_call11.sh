#!/bin/bash
mkdir -p gpurun_out
echo "== bench 2 gpus"; timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 2>gpurun_out/c11_bench2.err | tee gpurun_out/c11_bench2.json | cut -c1-1200
tail -2 gpurun_out/c11_bench2.err
echo "== probe"; timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/symm_probe.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -14
