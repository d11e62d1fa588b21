#!/bin/bash
# Round 2, GPU call 5 (one GPU): the deferred SH schedule -- parity test, then A/B of the bench.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_parity_r2_gpu.py tests/test_parity_gpu.py -m gpu -q -x > $O/r2c5_tests.txt 2>&1
tail -8 $O/r2c5_tests.txt
timeout 200 python bench.py --steps 100 --warmup 10 --quick --no-defer-sh > $O/r2c5_bench_nodefer.json 2> $O/r2c5_bench_nodefer.err
for c in 1 2 3 4; do
  timeout 200 python bench.py --steps 100 --warmup 10 --quick --defer-ctas $c > $O/r2c5_bench_defer$c.json 2> $O/r2c5_bench_defer$c.err
done
timeout 200 python bench.py --config c3 --steps 50 --warmup 5 --quick > $O/r2c5_bench_c3_defer.json 2> $O/r2c5_bench_c3_defer.err
for f in $O/r2c5_bench_*.err; do echo "== $f"; tail -3 $f; done
for f in $O/r2c5_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('kernels_ms'))"; done
