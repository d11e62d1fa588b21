#!/bin/bash
# Round 2, GPU call 14 (one GPU): `ncu --set full` of the kernels round 1 left without evidence.  The report stays on the box
# (it is far larger than what gpurun brings back); the raw page is exported as CSV.
set -x
mkdir -p gpurun_out
O=gpurun_out
K='regex:adam_kernel|l1_kernel|ssim_stats_kernel|ssim_grad_kernel|color_kernel|adam_rest_kernel|gather_rows_kernel|plan_|adc_flags|morton|bbox|zero_rows|gather_u32|slice4d_|filter3d_kernel|relocation_kernel|add_noise_kernel|blend_forward_kernel|blend_backward_kernel|preprocess_kernel|preprocess_backward_kernel'
timeout 620 ncu --set full --clock-control none --profile-from-start off -k "$K" -c 90 -f -o /tmp/r2_new_kernels python profiles/evidence_run.py > $O/r2c14_ncu_full.log 2>&1
tail -4 $O/r2c14_ncu_full.log
ls -la /tmp/r2_new_kernels.ncu-rep
ncu -i /tmp/r2_new_kernels.ncu-rep --page raw --csv > $O/r2_new_kernels_raw.csv 2> $O/r2c14_export.err
ls -la $O/r2_new_kernels_raw.csv
python profiles/summarize_ncu_r2.py $O/r2_new_kernels_raw.csv > $O/r2_kernels_ncu.md 2> $O/r2c14_summary.err; tail -3 $O/r2c14_summary.err; head -30 $O/r2_kernels_ncu.md | cut -c1-260
