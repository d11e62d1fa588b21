#!/bin/bash
# Round 2, GPU call 13 (one GPU): end-of-round confirmation of everything that became default + the ncu evidence set.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2c13_gpu_tests.txt 2>&1
tail -6 $O/r2c13_gpu_tests.txt
timeout 120 python __graft_entry__.py smoke > $O/r2c13_smoke.txt 2>&1; tail -2 $O/r2c13_smoke.txt
timeout 400 python bench.py --steps 100 --warmup 10 > $O/r2c13_bench.json 2> $O/r2c13_bench.err
timeout 300 python bench.py --steps 100 --warmup 10 --impl reference > $O/r2c13_bench_ref.json 2> $O/r2c13_bench_ref.err
timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o $O/r2_all_kernels python profiles/evidence_run.py > $O/r2c13_ncu_full.log 2>&1
tail -3 $O/r2c13_ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_mine.csv python bench.py --steps 2 --warmup 3 --quick > $O/r2c13_ncu_launches.log 2>&1
ls -la $O/r2_all_kernels.ncu-rep $O/r2_launches_mine.csv
for f in $O/r2c13_bench.json $O/r2c13_bench_ref.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dropin_unfused'), d.get('render_ms_1080p'))"; done
