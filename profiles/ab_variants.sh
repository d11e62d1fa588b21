#!/bin/bash
# A/B harness: run a micro-benchmark (default: the sort) for every experiment variant of libfgs_b200.so
SCRIPT=${1:-profiles/ab_sort.py}
for lib in faster-gaussian-splatting_b200/fgs_b200/lib/var/*.so; do
  echo "== $(basename $lib)"
  FGS_B200_LIB=$lib timeout 200 python $SCRIPT 2>&1 | tail -4
done
