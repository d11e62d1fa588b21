#!/bin/bash
# end-of-session validation: GPU tests, smoke, both bench arms, launch list and ncu details of the hot kernels
mkdir -p gpurun_out
echo "== gpu tests"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench mine"; timeout 500 python bench.py 2>gpurun_out/f_bench_mine.err | tee gpurun_out/f_bench_mine.json | cut -c1-300
echo "== bench reference"; timeout 400 python bench.py --impl reference 2>gpurun_out/f_bench_ref.err | tee gpurun_out/f_bench_ref.json | cut -c1-700
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/f_launches_mine.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; wc -l gpurun_out/f_launches_mine.csv
echo "== ncu full"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:"blend_|preprocess_backward|preprocess_kernel|onesweep|ranges_gather|create_instances|offsets_scan" -s 80 -c 16 -f -o gpurun_out/f_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu.log 2>&1; tail -1 gpurun_out/f_ncu.log
