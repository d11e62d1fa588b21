#!/bin/bash
# Round 2, GPU call 1 (one GPU): what round 1 wrote but never measured.
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2c1_smi.txt
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/r2c1_bench_default.json 2> $O/r2c1_bench_default.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --fused-loss > $O/r2c1_bench_fusedloss.json 2> $O/r2c1_bench_fusedloss.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --overlap-adam > $O/r2c1_bench_overlap.json 2> $O/r2c1_bench_overlap.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --overlap-adam --fused-loss > $O/r2c1_bench_overlap_fusedloss.json 2> $O/r2c1_bench_overlap_fusedloss.err
timeout 300 python bench.py --steps 100 --warmup 10 --impl reference > $O/r2c1_bench_ref.json 2> $O/r2c1_bench_ref.err
timeout 300 python profiles/bench_widen.py > $O/r2c1_widen_bench.json 2> $O/r2c1_widen_bench.err
timeout 300 python profiles/bench_fourd.py > $O/r2c1_fourd_bench.json 2> $O/r2c1_fourd_bench.err
timeout 600 python -m pytest tests/test_widen_training.py tests/test_widen_overlap_gpu.py -m gpu -q -s > $O/r2c1_training_tests.txt 2>&1
tail -5 $O/r2c1_training_tests.txt
