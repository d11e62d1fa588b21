#!/bin/bash
# Round 2, GPU call 18 (one GPU): combinations of the two blend-backward variants that won call 17.
mkdir -p gpurun_out
O=gpurun_out
run() { FGS_B200_LIB=$2 timeout 120 python bench.py --steps 100 --warmup 10 --quick > $O/r2c18_$1.json 2> $O/r2c18_$1.err; python -c "
import json
d=json.loads(open('$O/r2c18_$1.json').read().strip().splitlines()[-1]); k=d['kernels_ms']; print('$1', d['value'], d['ms_per_step'], 'fwd', k['blend_forward'], 'bwd', k['blend_backward'])"; }
V=$PWD/faster-gaussian-splatting_b200/fgs_b200/lib/var
run base ""
run hitlist_bwd_72 $V/libfgs_hb72.so
run hitlist_bwd_76 $V/libfgs_hb76.so
run hitlist_bwd_80 $V/libfgs_hb80.so
run hitlist_both_72 $V/libfgs_hfb72.so
run maxreg76 $V/libfgs_b76.so
