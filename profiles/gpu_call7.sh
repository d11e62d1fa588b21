#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== bench"; timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/c7_bench.err | tee gpurun_out/c7_bench.json
