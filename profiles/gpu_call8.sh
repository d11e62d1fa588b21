#!/bin/bash
mkdir -p gpurun_out
echo "== bench diag"; FGS_BENCH_DIAG=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/c8_bench.err | tee gpurun_out/c8_bench.json | cut -c1-400; grep diag gpurun_out/c8_bench.err
