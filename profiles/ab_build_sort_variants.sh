#!/bin/bash
# builds experiment variants of libfgs_b200.so next to the product library (git-ignored, travel to the GPU box)
set -e
cd "$(dirname "$0")/../faster-gaussian-splatting_b200"
rm -rf fgs_b200/lib/var; mkdir -p fgs_b200/lib/var
build() { name=$1; shift; python build.py --out fgs_b200/lib/var/$name.so "$@" > /dev/null; echo built $name; }
build w2 -DFGS_SORT_WINDOW=2
build w1 -DFGS_SORT_WINDOW=1
build w8 -DFGS_SORT_WINDOW=8
build cap2 -DFGS_SORT_WAVE_CAP=2
build cap2_i16 -DFGS_SORT_WAVE_CAP=2 -DFGS_SORT_ITEMS=16 -DFGS_SORT_MINBLOCKS=2
