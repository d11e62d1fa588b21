#!/bin/bash
# Round 2, GPU call 17 (one GPU): A/B of build variants of the blend kernels (VERDICT r1 item 5d and the XU-pipe idea).
mkdir -p gpurun_out
O=gpurun_out
run() { FGS_B200_LIB=$2 timeout 120 python bench.py --steps 100 --warmup 10 --quick > $O/r2c17_$1.json 2> $O/r2c17_$1.err; python -c "
import json
d=json.loads(open('$O/r2c17_$1.json').read().strip().splitlines()[-1]); k=d['kernels_ms']; print('$1', d['value'], d['ms_per_step'], 'fwd', k['blend_forward'], 'bwd', k['blend_backward'])"; }
V=$PWD/faster-gaussian-splatting_b200/fgs_b200/lib/var
run base ""
run hitlist $V/libfgs_hitlist.so
run bwdreg72 $V/libfgs_bwdreg72.so
run bwdreg96 $V/libfgs_bwdreg96.so
run bwdreg128 $V/libfgs_bwdreg128.so
run base2 ""
