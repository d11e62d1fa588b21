#!/bin/bash
# Round 2, GPU call 7 (two GPUs): compact PeerAdam (rows kernel) with and without the overlapped sh_rest broadcast.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_peer_adam.py -m gpu -x -q -s > $O/r2c7_peer_adam_test.txt 2>&1
tail -22 $O/r2c7_peer_adam_test.txt
for mode in peer-rows peer-rows-sync peer; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 60 --warmup 10 --quick --dp-mode $mode > $O/r2c7_bench_2gpu_$mode.json 2> $O/r2c7_bench_2gpu_$mode.err
  grep -v "OMP_NUM\|\*\*\*\*" $O/r2c7_bench_2gpu_$mode.err | tail -5
done
# single GPU: the deferred SH schedule with the colour kernel where it belongs
timeout 200 python bench.py --steps 100 --warmup 10 --quick --no-defer-sh > $O/r2c7_bench_nodefer.json 2> $O/r2c7_bench_nodefer.err
timeout 200 python bench.py --steps 100 --warmup 10 --quick --defer-ctas 2 > $O/r2c7_bench_defer2.json 2> $O/r2c7_bench_defer2.err
timeout 200 python bench.py --steps 100 --warmup 10 --quick --defer-ctas 1 > $O/r2c7_bench_defer1.json 2> $O/r2c7_bench_defer1.err
python profiles/defer_timeline.py 2 > $O/r2c7_defer_timeline2.txt 2>&1
for f in $O/r2c7_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('step_end_ms'), d.get('exposed_outside_rasterizer_ms'), d.get('kernels_ms'))"; done
