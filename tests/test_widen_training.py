"""End-to-end mini training on the GPU: the loop of /root/reference/Trainer.py:170-201 (render -> loss -> backward ->
FusedAdam.step) plus the every-N-iterations steps around it (adaptive density control Model.py:312-366, Morton
reordering :459-463), on a small synthetic multi-view problem whose targets are renders of a ground-truth scene.

What is compared (BASELINE.json: "PSNR vs ref"): the PSNR reached after the same number of iterations by
  * this repo's drop-in path vs the compiled reference backend (baseline/_ref), same loop, same torch L1 loss;
  * the fused backward+Adam path vs the drop-in path;
  * the fused loss kernel (csrc/loss.cu) vs torch's L1;
and that a run with L1 + DSSIM, densification and reordering in the loop stays healthy.  north_star asks for "PSNR equal
to 2 dp": measured on the B200 (round 2) all four runs -- drop-in, fused backward+Adam, fused loss kernel, compiled
reference -- reach 37.46 dB.  Float atomics make every run of either backend slightly different and 200 Adam steps carry
that along, so the test measures the reference's own run-to-run spread (two runs) and asserts
|PSNR difference| <= PSNR_TOL + 3 x spread, with PSNR_TOL = 0.01 dB (= 2 dp)."""
import math

import pytest
import torch

from tests import util

NAMES = ('means', 'scales', 'rotations', 'opacities', 'sh_coefficients_0', 'sh_coefficients_rest')
LRS = dict(means=1.6e-4, scales=5e-3, rotations=1e-3, opacities=2.5e-2, sh_coefficients_0=2.5e-3, sh_coefficients_rest=1.25e-4)
N, W, H, VIEWS, ITERS = 6000, 192, 144, 4, 200
PSNR_TOL = 0.01  # dB: north_star's "equal to 2 dp"


def _problem(device):
    import fgs_b200
    from fgs_b200 import synthetic as syn
    gt = syn.make_scene(N, 0, width=W).to(device)
    cams = [syn.make_camera(k, W, H, n_views=VIEWS) for k in range(VIEWS)]
    bg = torch.zeros(3, device=device)
    settings = [util.settings_for(c, device=device) for c in cams]
    with torch.no_grad():
        targets = [fgs_b200.rasterize(*gt, fgs_b200.RasterizerSettings(*s), to_chw=True) for s in settings]
    g = torch.Generator().manual_seed(42)
    init = {k: t.clone() for k, t in zip(NAMES, gt)}
    init['means'] += 0.02 * torch.randn(N, 3, generator=g).to(device)
    init['sh_coefficients_0'] += 0.3 * torch.randn(N, 1, 3, generator=g).to(device)
    init['opacities'] -= 1.0
    return settings, targets, init, bg


def _psnr(be, params, settings, targets):
    mse = 0.0
    with torch.no_grad():
        for s, t in zip(settings, targets):
            img = be.rasterize(*[params[k].detach() for k in NAMES], be.RasterizerSettings(*s), to_chw=True)
            mse += float(((img - t) ** 2).mean())
    return 10.0 * math.log10(1.0 / max(mse / len(targets), 1e-12))


def _train(be, init, settings, targets, iters=ITERS, fused=False, loss_fn=None, densify_at=(), morton_at=()):
    import fgs_b200
    params = {k: torch.nn.Parameter(v.clone().contiguous()) for k, v in init.items()}
    opt = be.FusedAdam([{'params': [params[k]], 'lr': LRS[k], 'name': k} for k in NAMES], lr=0.0, eps=1e-15)
    dev = params['means'].device
    track = bool(densify_at)
    info = torch.zeros(2, params['means'].shape[0], device=dev) if track else torch.empty(0)
    for it in range(iters):
        k = it % len(settings)
        rs = be.RasterizerSettings(*settings[k])
        args = [params[n] for n in NAMES]
        img = be.diff_rasterize_fused(*args, info, rs, opt) if fused else be.diff_rasterize(*args, info, rs)
        loss = loss_fn(img, targets[k]) if loss_fn is not None else (img - targets[k]).abs().mean()
        loss.backward()
        opt.step()
        opt.zero_grad()
        if it + 1 in densify_at:
            from fgs_b200 import densification as D
            params = D.adaptive_density_control(opt, info, 0.0002, 0.005, False, 0.01, 4.0)
            info = torch.zeros(2, params['means'].shape[0], device=dev)
        if it + 1 in morton_at:
            from fgs_b200 import densification as D
            params, (info_t,) = D.sort_param_groups(opt, D.morton_order(params['means'].detach()), extras=[info.t().contiguous()])
            info = info_t.t().contiguous()
    torch.cuda.synchronize()
    assert all(torch.isfinite(p).all() for p in params.values())
    return params


@pytest.mark.gpu
def test_training_psnr_parity(cuda_device):
    import fgs_b200
    from fgs_b200 import loss as L
    settings, targets, init, _ = _problem(cuda_device)
    psnr0 = _psnr(fgs_b200, init, settings, targets)
    p_drop = _psnr(fgs_b200, _train(fgs_b200, init, settings, targets), settings, targets)
    p_fused = _psnr(fgs_b200, _train(fgs_b200, init, settings, targets, fused=True), settings, targets)
    p_loss = _psnr(fgs_b200, _train(fgs_b200, init, settings, targets, loss_fn=lambda a, b: L.l1_dssim_loss(a, b, 1.0, 0.0)),
                   settings, targets)
    print(f'PSNR initial {psnr0:.2f} dB; drop-in {p_drop:.2f}, fused backward+Adam {p_fused:.2f}, fused L1 kernel {p_loss:.2f}')
    assert p_drop >= psnr0 + 1.0, (psnr0, p_drop)  # the loop trains
    ref = util.load_reference()
    spread = 0.0
    if ref is not None:
        p_ref = _psnr(fgs_b200, _train(ref, init, settings, targets), settings, targets)
        p_ref2 = _psnr(fgs_b200, _train(ref, init, settings, targets), settings, targets)
        spread = abs(p_ref - p_ref2)
        print(f'PSNR reference backend {p_ref:.3f} / {p_ref2:.3f} dB (run-to-run spread {spread:.4f}); this repo {p_drop:.3f} / fused {p_fused:.3f} / fused loss {p_loss:.3f}')
        assert abs(p_ref - p_drop) <= PSNR_TOL + 3 * spread, (p_ref, p_drop, spread)
    tol = PSNR_TOL + 3 * spread
    assert abs(p_fused - p_drop) <= tol, (p_fused, p_drop)
    assert abs(p_loss - p_drop) <= tol, (p_loss, p_drop)


@pytest.mark.gpu
def test_training_with_dssim_densification_and_reordering(cuda_device):
    import fgs_b200
    from fgs_b200 import loss as L
    settings, targets, init, _ = _problem(cuda_device)
    psnr0 = _psnr(fgs_b200, init, settings, targets)
    params = _train(fgs_b200, init, settings, targets, loss_fn=lambda a, b: L.l1_dssim_loss(a, b, 0.8, 0.2),
                    densify_at=(100,), morton_at=(150,))
    n_new = params['means'].shape[0]
    p = _psnr(fgs_b200, params, settings, targets)
    print(f'L1 + DSSIM with densification at 100 and Morton reordering at 150: N {N} -> {n_new}, PSNR {psnr0:.2f} -> {p:.2f} dB')
    # splitting replaces well-placed Gaussians by random samples of them, so no PSNR gain is asserted 50 iterations
    # after a densification round: only that the loop stays healthy on the new set of Gaussians
    assert n_new != N and math.isfinite(p) and p >= 10.0, (n_new, psnr0, p)
