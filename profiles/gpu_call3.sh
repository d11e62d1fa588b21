#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== ab sort"; timeout 300 python profiles/ab_sort.py faster-gaussian-splatting_b200/fgs_b200/lib/libfgs_b200.so faster-gaussian-splatting_b200/fgs_b200/lib/var/*.so 2>&1 | tee gpurun_out/c3_ab_sort.txt
echo "== bench"; timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/c3_bench.err | tee gpurun_out/c3_bench.json
echo "== ncu blend + binning"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blend_|create_instances|ranges_gather|preprocess_kernel" -s 40 -c 5 -f -o gpurun_out/c3_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c3_ncu.log 2>&1; tail -2 gpurun_out/c3_ncu.log
