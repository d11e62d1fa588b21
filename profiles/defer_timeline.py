#!/usr/bin/env python
"""Timeline of two consecutive training steps with the deferred SH schedule (CUDA events of the library's profiler, both
streams): where does the side-stream sh_rest update run relative to the next forward's kernels?

    python profiles/defer_timeline.py [ctas_per_sm] > gpurun_out/r2_defer_timeline.txt
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'faster-gaussian-splatting_b200')]


def main():
    import fgs_b200
    from fgs_b200 import _capi, loss as loss_mod, synthetic as syn
    from tests import util
    ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    _capi.load().fgs_debug_set_deferred_ctas_per_sm(ctas)
    n, w, h = 1_000_000, 1920, 1080
    sc = syn.make_scene(n, 0).to('cuda')
    lrs = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]
    params = [p.clone().requires_grad_(True) for p in sc]
    opt = fgs_b200.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs)], lr=0.0, eps=1e-15)
    rs = [fgs_b200.RasterizerSettings(*util.settings_for(syn.make_camera(k, w, h), device='cuda')) for k in range(8)]
    gt = [(syn.make_target(w, h, seed=k) * 255).round().to(torch.uint8).cuda() for k in range(8)]

    def step(i):
        img = fgs_b200.diff_rasterize_fused(*params, torch.empty(0), rs[i % 8], opt, defer_sh=True)
        terms, g = loss_mod.l1_dssim_loss_and_grad(img.detach(), gt[i % 8], 1.0, 0.0)
        img.backward(g)
        opt.step()
        opt.zero_grad()

    for i in range(20):
        step(i)
    torch.cuda.synchronize()
    _capi.profile_enable(True)
    for i in range(3):
        step(i)
    fgs_b200.join_deferred()
    torch.cuda.synchronize()
    _capi.profile_read()
    _capi.profile_enable(False)
    print(f'deferred update kernel: {ctas} CTA(s) per SM;  name, start (ms since the first kernel), end, duration')
    for name, t0, ms in _capi._last_starts:
        print(f'{name:40s} {t0:8.4f} {t0 + ms:8.4f} {ms:8.4f}')


if __name__ == '__main__':
    main()
