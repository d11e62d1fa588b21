#!/bin/bash
# Round 2, GPU call 16 (two GPUs): every way of ending a view-parallel step at HEAD (incl. the NVLS small-tensor option).
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_peer_adam.py -m gpu -x -q -s > $O/r2c16_peer_adam_test.txt 2>&1
grep "rank 0" $O/r2c16_peer_adam_test.txt; tail -3 $O/r2c16_peer_adam_test.txt
for mc in off on; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 60 --warmup 10 --quick --small-mc $mc > $O/r2c16_bench_2gpu_smallmc_$mc.json 2> $O/r2c16_bench_2gpu_smallmc_$mc.err
done
for f in $O/r2c16_bench_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('step_end_ms'))
    for r in d.get('per_rank', []): print('    ', r)
except Exception as e: print('$f', 'FAILED', e)"; done
for f in $O/r2c16_bench_*.err; do echo "== $f"; grep -v "OMP_NUM\|\*\*\*\*\|^$\|NCCL version" $f | tail -6; done
