"""CPU tests of the oracle: pinned against golden vectors dumped from the COMPILED REFERENCE on a B200
(tests/golden/*.npz, made by tests/golden/make_golden.py through baseline/_ref), plus internal
consistency checks (fp32 vs fp64, window rendering, Adam restatement vs torch.optim.Adam)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from tests import util

GOLDEN = Path(__file__).resolve().parent / 'golden'
NAMES = ['means', 'scales', 'rotations', 'opacities', 'sh0', 'sh_rest']


def _case(name):
    from fgs_b200 import synthetic as syn
    from oracle import torch_oracle as O
    z = np.load(GOLDEN / f'{name}.npz')
    n, w, h, seed, bases, aa = [int(v) for v in z['meta']]
    sc = syn.make_scene(n, seed)
    cam = syn.make_camera(seed, w, h)
    s = O.Settings(cam.w2c, cam.position, torch.from_numpy(z['bg']), bases, w, h, cam.focal_x, cam.focal_y,
                   cam.center_x, cam.center_y, cam.near_plane, cam.far_plane, bool(aa))
    gi = syn.make_grad_image(w, h, seed)
    return z, sc, s, gi


@pytest.mark.parametrize('name', ['c1_deg3', 'c1_deg1', 'c1_deg0', 'c1_aa', 'ragged'])
def test_oracle_matches_reference_golden(name):
    if not (GOLDEN / f'{name}.npz').exists():
        pytest.skip('golden fixture missing: oracle parity unpinned')
    from oracle import torch_oracle as O
    z, sc, s, gi = _case(name)
    dens = torch.zeros(2, sc.means.shape[0], dtype=torch.float64)
    out, grads = O.render_and_grads(sc.params(), s, gi, dtype=torch.float64, densification_info=dens)
    ref_img = torch.from_numpy(z['image']).double()
    l1 = (out.image - ref_img).abs().mean().item()
    assert l1 <= 1e-4, f'image mean-L1 vs reference {l1}'
    assert O.psnr(out.image.clamp(0, 1), ref_img.clamp(0, 1)) >= 60.0
    # inference path: clamped, HWC
    inf = torch.from_numpy(z['inference_hwc']).double().permute(2, 0, 1)
    assert (out.image.clamp(0, 1) - inf).abs().mean().item() <= 1e-4
    for k, g in zip(NAMES, grads):
        rg = torch.from_numpy(z[f'grad_{k}']).double()
        e = util.rel_err(g, rg)
        assert e <= 5e-3, f'{name}/{k}: |oracle - reference|_1 / |reference|_1 = {e}'
        vis = out.proj.visible
        assert float(rg[~vis].abs().sum()) == 0.0, 'reference rows of invisible Gaussians must be zero'
    rd = torch.from_numpy(z['densification_info']).double()
    assert torch.equal(rd[0], dens[0])
    assert util.rel_err(dens[1], rd[1]) <= 5e-3


def test_oracle_adam_matches_reference_golden():
    if not (GOLDEN / 'c1_deg3.npz').exists():
        pytest.skip('golden fixture missing')
    from oracle import torch_oracle as O
    z, sc, s, gi = _case('c1_deg3')
    lrs = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]
    for k, p0, lr in zip(NAMES, sc.params(), lrs):
        g = torch.from_numpy(z[f'grad_{k}'])
        p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        for step in (1, 2, 3):
            O.adam_step(p, g, m, v, step, lr)
        ref = torch.from_numpy(z[f'adam3_{k}'])
        assert util.rel_err(p - p0, ref - p0) <= 1e-3, k


def test_oracle_adam_is_torch_adam():
    from oracle import torch_oracle as O
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(257, 3, generator=g)
    a = p0.clone()
    b = p0.clone().requires_grad_(True)
    m, v = torch.zeros_like(a), torch.zeros_like(a)
    opt = torch.optim.Adam([b], lr=0.01, eps=1e-15)
    for step in (1, 2, 3, 4):
        gr = torch.randn(257, 3, generator=g)
        O.adam_step(a, gr, m, v, step, 0.01)
        b.grad = gr.clone()
        opt.step()
    assert torch.allclose(a, b.detach(), rtol=1e-4, atol=1e-6)


def test_oracle_window_equals_crop_of_full_render():
    from fgs_b200 import synthetic as syn
    from oracle import torch_oracle as O
    sc = syn.make_scene(400, 5)
    cam = syn.make_camera(5, 96, 72)
    s = O.Settings(cam.w2c.double(), cam.position.double(), torch.tensor([0.3, 0.2, 0.1]), 16, 96, 72, cam.focal_x,
                   cam.focal_y, cam.center_x, cam.center_y, cam.near_plane, cam.far_plane, False)
    p64 = [p.double() for p in sc.params()]
    with torch.no_grad():
        full = O.render(*p64, s).image
        win = O.render(*p64, s, window=(16, 24, 64, 60)).image
    assert torch.allclose(win, full[:, 24:60, 16:64], atol=1e-12)


def test_oracle_fp32_close_to_fp64():
    from fgs_b200 import synthetic as syn
    from oracle import torch_oracle as O
    sc = syn.make_scene(300, 7)
    cam = syn.make_camera(7, 80, 60)
    s = O.Settings(cam.w2c, cam.position, torch.zeros(3), 16, 80, 60, cam.focal_x, cam.focal_y, cam.center_x,
                   cam.center_y, cam.near_plane, cam.far_plane, False)
    gi = syn.make_grad_image(80, 60, 7)
    o64, g64 = O.render_and_grads(sc.params(), s, gi, dtype=torch.float64)
    o32, g32 = O.render_and_grads(sc.params(), s, gi, dtype=torch.float32)
    assert (o64.image - o32.image.double()).abs().mean().item() <= 1e-5
    for a, b in zip(g32, g64):
        assert util.rel_err(a, b) <= 2e-3


def test_oracle_pruning_scores_matches_reference_golden():
    f = GOLDEN / 'pruning_scores.npz'
    if not f.exists():
        pytest.skip('golden fixture missing')
    from fgs_b200 import synthetic as syn
    from oracle import torch_oracle as O
    z = np.load(f)
    n, w, h, seed = [int(v) for v in z['meta']]
    sc = syn.make_scene(n, seed)
    cam = syn.make_camera(seed, w, h)
    s = O.Settings(cam.w2c, cam.position, torch.from_numpy(z['bg']), 16, w, h, cam.focal_x, cam.focal_y, cam.center_x,
                   cam.center_y, cam.near_plane, cam.far_plane, False)
    got = O.pruning_scores(*sc, s)
    assert util.rel_err(got, torch.from_numpy(z['scores']).double()) <= 5e-3


def test_binning_oracle_is_consistent_with_preprocess_and_dense_blend():
    """The integer restatement of the binning stage (oracle.binning): instance count == sum of the exact tile
    counts of preprocess, every tile's list is front-to-back, and it is complete w.r.t. the dense blend: every
    (pixel, Gaussian) pair with alpha >= 1/255 has the Gaussian in the list of the pixel's tile,
    so the tiled reference algorithm and the dense oracle blend the same pairs."""
    from fgs_b200 import synthetic as syn
    from oracle import torch_oracle as O
    n, w, h = 600, 96, 60
    sc = syn.make_scene(n, 3)
    cam = syn.make_camera(3, w, h)
    s = O.Settings(cam.w2c, cam.position, torch.zeros(3), 16, w, h, cam.focal_x, cam.focal_y, cam.center_x,
                   cam.center_y, cam.near_plane, cam.far_plane, False)
    params = [p.double() for p in sc]
    proj = O.preprocess(*params, s)
    bn = O.binning(proj, s)
    grid_w, grid_h = (w + 15) // 16, (h + 11) // 12
    assert bn.tile_ranges.shape == (grid_w * grid_h, 2)
    assert int(bn.instances.numel()) == int(proj.n_touched_tiles[proj.visible].sum()) > 0
    assert int(bn.tile_ranges[-1, 1]) == bn.instances.numel() and int(bn.tile_ranges[0, 0]) == 0
    assert torch.equal(bn.tile_ranges[1:, 0], bn.tile_ranges[:-1, 1])          # contiguous, tile-major
    assert torch.equal(bn.tile_ids, torch.sort(bn.tile_ids).values)
    member = torch.zeros(grid_w * grid_h, n, dtype=torch.bool)
    for t in range(grid_w * grid_h):
        a, b = (int(v) for v in bn.tile_ranges[t])
        lst = bn.instances[a:b]
        assert bool(proj.visible[lst].all())
        d = proj.depth[lst]
        assert bool((d[1:] >= d[:-1]).all()), 'tile list not front-to-back'
        assert lst.unique().numel() == lst.numel(), 'a Gaussian twice in one tile list'
        member[t, lst] = True
    # dense alphas of all visible Gaussians on all pixels (kernels_forward.cuh:460-465)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing='ij')
    px, py = xs.reshape(-1) + 0.5, ys.reshape(-1) + 0.5
    vis = proj.visible.nonzero(as_tuple=True)[0]
    dx = proj.mean2d[vis, 0:1] - px[None]
    dy = proj.mean2d[vis, 1:2] - py[None]
    c = proj.conic[vis]
    q = -0.5 * (c[:, 0:1] * dx * dx + c[:, 2:3] * dy * dy) - c[:, 1:2] * dx * dy
    alpha = proj.opacity[vis, None] * torch.exp(torch.clamp(q, max=0.0))
    tile_of_pixel = ((ys.reshape(-1) // 12) * grid_w + xs.reshape(-1) // 16).long()
    need = alpha >= (1.0 / 255.0) * (1.0 + 1e-9)  # strictly above the threshold: float edge cases aside
    listed = member[tile_of_pixel][:, vis].T       # (V, P): Gaussian in the list of the pixel's tile
    assert int(need.sum()) > 1000
    assert bool(listed[need].all()), 'a contributing (pixel, Gaussian) pair is missing from the tile lists'
    # and the lists are tight: most listed pairs' tiles do contain a contributing pixel
    assert listed.any(dim=1).all()


def test_binning_oracle_empty_scene():
    from fgs_b200 import synthetic as syn
    from oracle import torch_oracle as O
    sc = syn.make_scene(10, 0)
    cam = syn.make_camera(0, 64, 48)
    s = O.Settings(cam.w2c, cam.position, torch.zeros(3), 16, 64, 48, cam.focal_x, cam.focal_y, cam.center_x,
                   cam.center_y, 1e6, 1e7, False)  # everything in front of the near plane
    proj = O.preprocess(*sc, s)
    bn = O.binning(proj, s)
    assert bn.instances.numel() == 0 and int(bn.tile_ranges.sum()) == 0
