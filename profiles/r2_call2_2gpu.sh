#!/bin/bash
# Round 2, GPU call 2 (two GPUs): first hardware run of PeerAdam (peer + NVLS multicast) and ShardedAdam, then the
# view-parallel step with every way of ending it.
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2c2_topo.txt 2>&1
timeout 300 python -m pytest tests/test_peer_adam.py -m gpu -x -q -s > $O/r2c2_peer_adam_test.txt 2>&1
tail -12 $O/r2c2_peer_adam_test.txt
for mode in peer-mc peer sharded flat allreduce pipelined; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 60 --warmup 10 --quick --dp-mode $mode > $O/r2c2_bench_2gpu_$mode.json 2> $O/r2c2_bench_2gpu_$mode.err
  tail -c 600 $O/r2c2_bench_2gpu_$mode.json; tail -3 $O/r2c2_bench_2gpu_$mode.err
done
# the named configs[3]: 8-view batch split over the two ranks (strong scaling)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 30 --warmup 5 --quick --batch 8 > $O/r2c2_bench_2gpu_batch8.json 2> $O/r2c2_bench_2gpu_batch8.err
tail -3 $O/r2c2_bench_2gpu_batch8.err
