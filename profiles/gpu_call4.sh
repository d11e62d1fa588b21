#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== ab sort"; timeout 300 python profiles/ab_sort.py faster-gaussian-splatting_b200/fgs_b200/lib/libfgs_b200.so faster-gaussian-splatting_b200/fgs_b200/lib/var/*.so 2>&1 | head -18 | tee gpurun_out/c4_ab_sort.txt
echo "== bench"; timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/c4_bench.err | tee gpurun_out/c4_bench.json
