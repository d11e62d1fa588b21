#!/usr/bin/env python
"""BASELINE.json configs[4] (FasterGS4D shape: 500 k 4D Gaussians x 150 time steps, 1080p) for THIS repo only -- the
reference's FasterGS4D branch is not in the tree, so there is no reference arm.  Prints one JSON object: render ms per
time step (inference path, 150 steps) and training ms per iteration (time slice + forward + L1 + backward, gradients of
all seven tensors, random time per iteration), CUDA events, inputs resident.

    python profiles/bench_fourd.py > gpurun_out/fourd_bench.json
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'faster-gaussian-splatting_b200')]


def main():
    import fgs_b200
    from fgs_b200 import fourd as F4, synthetic as syn
    from tests import util
    from tests.test_widen_fourd import make_scene_4d
    assert torch.cuda.is_available(), 'needs a GPU'
    dev = torch.device('cuda', 0)
    n, w, h, steps = 500_000, 1920, 1080, 150
    params = [p.to(dev) for p in make_scene_4d(n, 0, w)]
    cams = [syn.make_camera(k, w, h) for k in range(8)]
    rs = [fgs_b200.RasterizerSettings(*util.settings_for(c, device=dev)) for c in cams]
    target = syn.make_target(w, h).to(dev)
    times = [i / (steps - 1) for i in range(steps)]
    for i in range(10):
        F4.rasterize_4d(*params, times[i], rs[i % 8], to_chw=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i, t in enumerate(times):
        F4.rasterize_4d(*params, t, rs[i % 8], to_chw=True)
    e1.record()
    torch.cuda.synchronize()
    render_ms = e0.elapsed_time(e1) / steps
    leaves = [p.clone().requires_grad_(True) for p in params]

    def train(i):
        img = F4.diff_rasterize_4d(*leaves, times[(37 * i) % steps], torch.empty(0), rs[i % 8])
        (img - target).abs().mean().backward()
        for p in leaves:
            p.grad = None
    for i in range(10):
        train(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(100):
        train(i)
    e1.record()
    torch.cuda.synchronize()
    train_ms = e0.elapsed_time(e1) / 100
    print(json.dumps({'workload': 'configs[4]: 500k 4D Gaussians, 1920x1080, 150 time steps', 'render_ms_per_timestep': round(render_ms, 4),
                      'train_ms_per_iteration_fwd_l1_bwd': round(train_ms, 4), 'reference': 'none (FasterGS4D branch not in the tree)'}))


if __name__ == '__main__':
    main()
