#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
echo "== probe"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/symm_probe.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -30
echo "== bench 2 gpus"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 5 2>gpurun_out/c10_bench2.err | tee gpurun_out/c10_bench2.json | cut -c1-900
tail -3 gpurun_out/c10_bench2.err
