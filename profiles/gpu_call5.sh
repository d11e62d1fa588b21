#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"create_instances|offsets_scan|blend_forward" -s 24 -c 3 -f -o gpurun_out/c5_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c5_ncu.log 2>&1; tail -2 gpurun_out/c5_ncu.log
