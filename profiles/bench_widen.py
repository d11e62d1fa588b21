#!/usr/bin/env python
"""Micro-benchmark of the widened rows on one GPU, at BASELINE.json's shapes:

  * loss:    L1 and L1 + DSSIM with gradient at 1920x1080 (csrc/loss.cu) next to the torch ops the reference's Loss.py
             issues (l1_loss, and an SSIM built from conv2d as a stand-in for NeRFICG's fused_dssim, which is not in
             the tree);
  * densify: prune 30 %, permutation (sort), Morton order, adaptive density control on 1 M Gaussians with Adam state
             (csrc/densify.cu) next to the same operations written with torch indexing ops (what Model.py +
             adam_utils do).

CUDA events on the current stream, 5 warm-ups, median of 20; inputs (>= 700 MB for the row moves) exceed L2.  Prints one
JSON object: milliseconds, algorithmic bytes, achieved GB/s and the fraction of the measured HBM peak.

    python profiles/bench_widen.py > gpurun_out/widen_bench.json
"""
from __future__ import annotations

import json
import math
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'faster-gaussian-splatting_b200'))
sys.path.insert(0, str(ROOT))

NAMES = ('means', 'scales', 'rotations', 'opacities', 'sh_coefficients_0', 'sh_coefficients_rest')


def peak_gbs():
    f = ROOT / 'MEASURED_PEAKS.json'
    try:
        return float(json.loads(f.read_text())['hbm_gbs'])
    except Exception:
        return 6650.0


def timeit(fn, setup=None, reps=20, warmup=5):
    ts = []
    for i in range(warmup + reps):
        if setup is not None:
            setup()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def row(ms, alg_bytes, peak):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    return {'ms': round(ms, 4), 'algorithmic_MB': round(alg_bytes / 1e6, 1), 'GBps': round(gbs, 1), 'frac_of_hbm_peak': round(gbs / peak, 3)}


def torch_ssim_loss(x, y, win):
    import torch.nn.functional as F
    c = x.shape[0]
    kh, kv = win.reshape(1, 1, 1, 11).expand(c, 1, 1, 11), win.reshape(1, 1, 11, 1).expand(c, 1, 11, 1)
    blur = lambda t: F.conv2d(F.conv2d(t[None], kh, padding=(0, 5), groups=c), kv, padding=(5, 0), groups=c)[0]
    mu1, mu2 = blur(x), blur(y)
    s11, s22, s12 = blur(x * x) - mu1 * mu1, blur(y * y) - mu2 * mu2, blur(x * y) - mu1 * mu2
    m = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s11 + s22 + 9e-4))
    return 1.0 - m.mean()


def bench_loss(dev, peak):
    from fgs_b200 import loss as L
    w, h = 1920, 1080
    n = 3 * w * h
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, h, w, generator=g).to(dev)
    y8 = (torch.rand(3, h, w, generator=g) * 255).round().to(torch.uint8).to(dev)
    yf = y8.float().div(255.0)
    k = torch.arange(11, dtype=torch.float32, device=dev) - 5
    win = torch.exp(-k * k / 4.5)
    win = win / win.sum()
    out = {}
    out['l1_fused_u8_target'] = row(timeit(lambda: L.l1_dssim_loss_and_grad(x, y8, 1.0, 0.0)), 4 * n + n + 4 * n, peak)
    out['l1_fused_f32_target'] = row(timeit(lambda: L.l1_dssim_loss_and_grad(x, yf, 1.0, 0.0)), 12 * n, peak)

    def torch_l1():
        xr = x.detach().requires_grad_(True)
        (xr - yf).abs().mean().backward()
    out['l1_torch'] = row(timeit(torch_l1), 12 * n, peak)
    # L1 + DSSIM: stats kernel (x, y in; 3 maps out) + grad kernel (3 maps, x, y in; grad out)
    out['l1_dssim_fused_u8_target'] = row(timeit(lambda: L.l1_dssim_loss_and_grad(x, y8, 0.8, 0.2)), (5 + 12 + 12 + 5 + 4) * n, peak)
    out['l1_dssim_fused_f32_target'] = row(timeit(lambda: L.l1_dssim_loss_and_grad(x, yf, 0.8, 0.2)), (8 + 12 + 12 + 8 + 4) * n, peak)

    def torch_l1_dssim():
        xr = x.detach().requires_grad_(True)
        (0.8 * (xr - yf).abs().mean() + 0.2 * torch_ssim_loss(xr, yf, win)).backward()
    out['l1_dssim_torch_conv2d'] = row(timeit(torch_l1_dssim), (8 + 12 + 12 + 8 + 4) * n, peak)
    return out


def make_optimizer(n, dev):
    import fgs_b200
    from fgs_b200 import synthetic as syn
    sc = syn.make_scene(n, 0)
    params = {k: torch.nn.Parameter(t.to(dev).contiguous()) for k, t in zip(NAMES, sc)}
    params['scales'].data[::2] += 2.5  # half of the Gaussians above the clone / split size threshold
    opt = fgs_b200.FusedAdam([{'params': [params[k]], 'lr': 1e-3, 'name': k} for k in NAMES], lr=0.0, eps=1e-15)
    for k in NAMES:
        opt.state[params[k]] = {'step': 1, 'exp_avg': torch.randn_like(params[k]) * 1e-3, 'exp_avg_sq': torch.rand_like(params[k]) * 1e-6}
    return opt


def torch_select(opt, index_or_mask):
    """prune / sort with torch indexing: what adam_utils.prune_param_groups / sort_param_groups do per group."""
    for group in opt.param_groups:
        p = group['params'][0]
        st = opt.state.pop(p)
        new_p = torch.nn.Parameter(p.data[index_or_mask].contiguous())
        opt.state[new_p] = {'step': st['step'], 'exp_avg': st['exp_avg'][index_or_mask].contiguous(),
                            'exp_avg_sq': st['exp_avg_sq'][index_or_mask].contiguous()}
        group['params'][0] = new_p


def bench_densify(dev, peak):
    from fgs_b200 import densification as D
    n = 1_000_000
    row_bytes = 59 * 4 * 3  # parameters + two moments per Gaussian
    out = {}
    g = torch.Generator().manual_seed(1)
    keep = (torch.rand(n, generator=g) < 0.7).to(dev)
    n_keep = int(keep.sum())
    perm = torch.randperm(n, generator=g).to(dev)
    state = {}

    def fresh():
        state['opt'] = make_optimizer(n, dev)

    out['prune_30pct_fused'] = row(timeit(lambda: D.prune_param_groups(state['opt'], keep), fresh, reps=8, warmup=2),
                                   n + n_keep * 2 * row_bytes, peak)
    out['prune_30pct_torch_indexing'] = row(timeit(lambda: torch_select(state['opt'], keep), fresh, reps=8, warmup=2),
                                            n + n_keep * 2 * row_bytes, peak)
    out['sort_permutation_fused'] = row(timeit(lambda: D.sort_param_groups(state['opt'], perm), fresh, reps=8, warmup=2),
                                        8 * n + n * 2 * row_bytes, peak)
    out['sort_permutation_torch_indexing'] = row(timeit(lambda: torch_select(state['opt'], perm), fresh, reps=8, warmup=2),
                                                 8 * n + n * 2 * row_bytes, peak)
    fresh()
    means = state['opt'].param_groups[0]['params'][0].detach()
    out['morton_order_fused'] = row(timeit(lambda: D.morton_order(means), reps=8, warmup=2), 12 * n + 6 * 16 * n + 8 * n, peak)
    info = torch.stack([torch.randint(0, 40, (n,), generator=g).float(), torch.rand(n, generator=g) * 0.012]).to(dev)

    def adc():
        D.adaptive_density_control(state['opt'], info, 0.0002, 0.005, True, 0.01, 4.0)
    ms = timeit(adc, fresh, reps=8, warmup=2)
    n_out = state['opt'].param_groups[0]['params'][0].shape[0]
    out['adaptive_density_control_fused'] = row(ms, 8 * n + 45 * n + n_out * (8 + 2 * row_bytes), peak)
    out['adaptive_density_control_fused']['n_out'] = n_out
    return out


def main():
    assert torch.cuda.is_available(), 'needs a GPU'
    dev = torch.device('cuda', 0)
    peak = peak_gbs()
    res = {'hbm_peak_GBps': peak, 'loss_1080p': bench_loss(dev, peak), 'densify_1M': bench_densify(dev, peak)}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
