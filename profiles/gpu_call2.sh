#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^onesweep_kernel" -s 8 -c 2 -f -o gpurun_out/c2_sort_u32 python profiles/ab_sort.py > gpurun_out/c2_ncu_u32.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^onesweep_kernel" -s 52 -c 2 -f -o gpurun_out/c2_sort_u16 python profiles/ab_sort.py > gpurun_out/c2_ncu_u16.log 2>&1
ls -la gpurun_out/
