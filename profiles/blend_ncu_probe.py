#!/usr/bin/env python
"""Two training steps at C2 for a metrics-only ncu pass over the blend kernels (instruction counts and pipe utilisation of the
final kernels, for the issue roofline in profiles/r2_results.md):

    ncu --clock-control none -k regex:blend_ -s 4 -c 2 --metrics <list> --csv --log-file gpurun_out/r2_blend_final.csv python profiles/blend_ncu_probe.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'faster-gaussian-splatting_b200')]
import fgs_b200  # noqa: E402
from fgs_b200 import synthetic as syn  # noqa: E402
from tests import util  # noqa: E402

sc = syn.make_scene(1_000_000, 0).to('cuda')
st = util.settings_for(syn.make_camera(0, 1920, 1080), device='cuda')
gt = syn.make_target(1920, 1080).cuda()
for _ in range(3):
    params = [p.clone().requires_grad_(True) for p in sc]
    img = fgs_b200.diff_rasterize(*params, torch.empty(0), fgs_b200.RasterizerSettings(*st))
    (img - gt).abs().mean().backward()
torch.cuda.synchronize()
print('done')
