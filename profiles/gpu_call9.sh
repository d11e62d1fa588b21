#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== bench"; FGS_BENCH_DIAG=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/c9_bench.err | tee gpurun_out/c9_bench.json; grep diag gpurun_out/c9_bench.err
