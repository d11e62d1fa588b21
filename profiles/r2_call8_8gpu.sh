#!/bin/bash
# Round 2, GPU call 8 (eight GPUs): the view-parallel step at 8 and 4 ranks, every way of ending it that matters.
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2c8_topo.txt 2>&1
run() {  # n_gpus mode extra...
  n=$1; mode=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 60 --warmup 10 --quick --dp-mode $mode "$@" > $O/r2c8_bench_${n}gpu_$mode$(echo "$@" | tr -d ' -').json 2> $O/r2c8_bench_${n}gpu_$mode$(echo "$@" | tr -d ' -').err
}
run 8 peer-rows
run 8 peer
run 8 peer-mc
run 8 sharded
run 4 peer-rows
run 4 peer-rows --batch 8
for f in $O/r2c8_bench_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('step_end_ms'), d.get('exposed_outside_rasterizer_ms'), d.get('kernels_ms_sum'))
except Exception as e: print('$f', 'FAILED', e)"; done
for f in $O/r2c8_bench_*.err; do echo "== $f"; grep -v "OMP_NUM\|\*\*\*\*\|^$" $f | tail -4; done
