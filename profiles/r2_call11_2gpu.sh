#!/bin/bash
# Round 2, GPU call 10 (two GPUs): the rows kernel rewritten for memory-level parallelism -- correctness, then phases.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_peer_adam.py -m gpu -x -q -s > $O/r2c11_peer_adam_test.txt 2>&1
grep "rank 0" $O/r2c11_peer_adam_test.txt; tail -3 $O/r2c11_peer_adam_test.txt
run() {  # n_gpus mode extra...
  n=$1; mode=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 60 --warmup 10 --quick --dp-mode $mode "$@" > $O/r2c11_bench_${n}gpu_$mode$(echo "$@" | tr -d ' -').json 2> $O/r2c11_bench_${n}gpu_$mode$(echo "$@" | tr -d ' -').err
}
run 2 peer-rows
run 2 peer-rows-sync
run 2 peer-rows-mono-p2p
run 2 auto --batch 8
for f in $O/r2c11_bench_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('step_end_ms'), d.get('exposed_outside_rasterizer_ms'), d.get('kernels_ms_sum'))
    for r in d.get('per_rank', []): print('    ', r)
except Exception as e: print('$f', 'FAILED', e)"; done
for f in $O/r2c11_bench_*.err; do echo "== $f"; grep -v "OMP_NUM\|\*\*\*\*\|^$\|NCCL version" $f | tail -12; done
