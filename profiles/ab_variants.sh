#!/bin/bash
# A/B harness: run the bench's per-kernel profile for every experiment variant of libfgs_b200.so
for lib in faster-gaussian-splatting_b200/fgs_b200/lib/var/*.so; do
  echo -n "$(basename $lib): "
  FGS_B200_LIB=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print(d['ms_per_step'], 'bwd', k['blend_backward'], 'fwd', k['blend_forward'])"
done
