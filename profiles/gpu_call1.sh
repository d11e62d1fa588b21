#!/bin/bash
# GPU call 1 of this session: correctness of the chunked onesweep sort + A/B of its variants + step profile
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
echo "== sort test"; timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "radix" 2>&1 | tail -5
echo "== full gpu suite"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== ab sort"; timeout 300 python profiles/ab_sort.py faster-gaussian-splatting_b200/fgs_b200/lib/libfgs_b200.so faster-gaussian-splatting_b200/fgs_b200/lib/var/*.so 2>&1 | tee gpurun_out/c1_ab_sort.txt
echo "== bench"; timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/c1_bench.err | tee gpurun_out/c1_bench.json
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"onesweep|offsets_scan|create_instances|radix_hist|ranges_gather" -c 60 --csv --log-file gpurun_out/c1_sort_launches.csv -s 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; tail -3 gpurun_out/c1_sort_launches.csv
