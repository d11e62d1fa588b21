#!/bin/bash
# Round 2, GPU call 3 (one GPU): parity noise floor, all GPU tests, the missing bench lines, blend counters.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python profiles/parity_table.py > $O/r2_parity.md 2> $O/r2_parity.err
tail -3 $O/r2_parity.err
timeout 900 python -m pytest tests -m gpu -q -s > $O/r2c3_gpu_tests.txt 2>&1
tail -25 $O/r2c3_gpu_tests.txt
FGS_B200_LIB=$PWD/faster-gaussian-splatting_b200/fgs_b200/lib/libfgs_b200_stats.so timeout 300 python profiles/blend_stats.py > $O/r2_blend_stats.json 2> $O/r2_blend_stats.err
timeout 400 python bench.py --steps 100 --warmup 10 > $O/r2c3_bench_c2.json 2> $O/r2c3_bench_c2.err
timeout 300 python bench.py --steps 100 --warmup 10 --impl reference > $O/r2c3_bench_c2_ref.json 2> $O/r2c3_bench_c2_ref.err
timeout 300 python bench.py --config c3 --steps 50 --warmup 5 --no-cpu-baseline > $O/r2c3_bench_c3.json 2> $O/r2c3_bench_c3.err
timeout 300 python bench.py --config c3 --steps 50 --warmup 5 --impl reference > $O/r2c3_bench_c3_ref.json 2> $O/r2c3_bench_c3_ref.err
timeout 300 python bench.py --config c5 --steps 100 --warmup 10 > $O/r2c3_bench_c5.json 2> $O/r2c3_bench_c5.err
timeout 300 python bench.py --steps 50 --warmup 5 --quick --yaml-loss > $O/r2c3_bench_yaml.json 2> $O/r2c3_bench_yaml.err
timeout 300 python bench.py --steps 50 --warmup 5 --quick --yaml-loss --impl reference > $O/r2c3_bench_yaml_ref.json 2> $O/r2c3_bench_yaml_ref.err
timeout 300 python bench.py --steps 20 --warmup 3 --quick --batch 8 > $O/r2c3_bench_batch8.json 2> $O/r2c3_bench_batch8.err
timeout 300 python bench.py --steps 20 --warmup 3 --quick --batch 8 --impl reference > $O/r2c3_bench_batch8_ref.json 2> $O/r2c3_bench_batch8_ref.err
for f in $O/r2c3_bench_*.err; do echo "== $f"; tail -3 $f; done
