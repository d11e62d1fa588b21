#!/bin/bash
# Round 2, GPU call 12 (eight GPUs): the split compact view-parallel step at 8 and 4 ranks, next to the NVLS one-kernel step.
set -x
mkdir -p gpurun_out
O=gpurun_out
run() {  # n_gpus mode extra...
  n=$1; mode=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 60 --warmup 10 --dp-mode $mode "$@" > $O/r2c12_bench_${n}gpu_$mode$(echo "$@" | tr -d ' -').json 2> $O/r2c12_bench_${n}gpu_$mode$(echo "$@" | tr -d ' -').err
}
run 8 auto
run 8 peer-rows-sync --quick
run 8 peer-mc --quick
run 4 auto --quick
run 4 auto --quick --batch 8
for f in $O/r2c12_bench_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('step_end_ms'), d.get('exposed_outside_rasterizer_ms'), d.get('kernels_ms_sum'))
    for r in d.get('per_rank', []): print('    ', r)
except Exception as e: print('$f', 'FAILED', e)"; done
for f in $O/r2c12_bench_*.err; do echo "== $f"; grep -v "OMP_NUM\|\*\*\*\*\|^$\|NCCL version" $f | tail -6; done
