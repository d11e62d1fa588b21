"""GPU parity, second part: the noise floor that sets the tolerances, the inference path against the reference's own
`rasterize()`, the reference's unmodified Python bindings running over this repo's `_C` (INTEGRATION.md option B), and
the per-device / per-shape host state.

Tolerances are north_star's: image mean-L1 <= 1e-4, gradients 1e-3 (per tensor, |a-b|_1/|b|_1).  Where the compiled
reference is available the floor is measured in the test itself: the reference against the fp64 oracle (its own
fast-math) and against itself run-to-run (float atomics); this repo has to be inside 1e-3 AND inside
max(3 x floor, 1e-4) of the oracle, so a regression that hides below north_star's bound but above the reference's own
accuracy still fails.  Measured values: profiles/r2_parity.md."""
import sys
import types
from pathlib import Path

import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
NAMES = ['means', 'scales', 'rotations', 'opacities', 'sh0', 'sh_rest']
GRAD_RTOL = 1e-3   # north_star: "gradients to rtol 1e-3"
IMAGE_L1 = 1e-4    # north_star: "images match to <= 1e-4 L1"


def _run(be, sc, st, gi):
    params = [p.clone().requires_grad_(True) for p in sc]
    img = be.diff_rasterize(*params, torch.empty(0), be.RasterizerSettings(*st))
    img.backward(gi)
    return img.detach(), [p.grad.detach() for p in params]


def _oracle(sc, st, gi):
    from oracle import torch_oracle as O
    cpu = [p.detach().cpu() for p in sc]
    s = O.Settings(*[t.cpu() if torch.is_tensor(t) else t for t in st])
    out, grads = O.render_and_grads(cpu, s, gi.cpu(), dtype=torch.float64)
    return out.image, grads


@pytest.mark.parametrize('seed,aa', [(0, False), (1, False), (2, True)])
def test_reference_vs_oracle_sets_the_scale(cuda_device, seed, aa):
    """C1: (i) reference twice, (ii) reference vs fp64 oracle, (iii) this repo vs fp64 oracle, (iv) this repo vs reference."""
    import fgs_b200
    from fgs_b200 import synthetic as syn
    sc = syn.make_scene(1000, seed).to('cuda')
    cam = syn.make_camera(seed, 128, 128)
    st = util.settings_for(cam, bg=(0.2, 0.3, 0.1), aa=aa, device='cuda')
    gi = syn.make_grad_image(128, 128, seed).cuda()
    m_img, m_g = _run(fgs_b200, sc, st, gi)
    o_img, o_g = _oracle(sc, st, gi)
    ref = util.load_reference()
    l1 = (m_img.cpu().double() - o_img).abs().mean().item()
    assert l1 <= IMAGE_L1, f'image mean-L1 vs fp64 oracle {l1}'
    floor = {n: 0.0 for n in NAMES}
    if ref is not None:
        r_img, r_g = _run(ref, sc, st, gi)
        r_img2, r_g2 = _run(ref, sc, st, gi)
        assert (m_img - r_img).abs().mean().item() <= IMAGE_L1
        for n, a, b, o in zip(NAMES, r_g, r_g2, o_g):
            floor[n] = max(util.rel_err(a.cpu(), o), util.rel_err(a, b))
    rows = []
    for k, n in enumerate(NAMES):
        e_or = util.rel_err(m_g[k].cpu(), o_g[k])
        e_ref = util.rel_err(m_g[k], r_g[k]) if ref is not None else float('nan')
        rows.append((n, floor[n], e_or, e_ref))
        assert e_or <= GRAD_RTOL, f'{n}: {e_or} vs fp64 oracle (north_star 1e-3)'
        if ref is not None:
            assert e_ref <= GRAD_RTOL, f'{n}: {e_ref} vs compiled reference (north_star 1e-3)'
            assert e_or <= max(3.0 * floor[n], 1e-4), f'{n}: {e_or} vs oracle, the reference itself is at {floor[n]}'
    print('\n' + '\n'.join(f'  {n:10s} reference floor {f:.1e}  mine vs oracle {a:.1e}  mine vs reference {b:.1e}' for n, f, a, b in rows))


def test_rasterize_matches_reference_rasterize(cuda_device):
    """The "render ms" path: rasterize() against the reference's rasterize() (inference.cu:11-226) for both layouts and
    both clamp settings, at C1 and at a 50 k / 480x270 scene with the depth complexity of the benchmark."""
    import fgs_b200
    from fgs_b200 import synthetic as syn
    from oracle import torch_oracle as O
    ref = util.load_reference()
    if ref is None:
        pytest.skip('baseline/_ref not present')
    for n, w, h, seed in ((1000, 128, 128, 3), (50_000, 480, 270, 0), (1_000_000, 1920, 1080, 0)):
        sc = syn.make_scene(n, seed, width=w if n != 1000 else 1920).to('cuda')
        cam = syn.make_camera(seed, w, h)
        st = util.settings_for(cam, bg=(0.7, 0.1, 0.4), device='cuda')
        for to_chw in (True, False):
            for clamp in (True, False):
                a = fgs_b200.rasterize(*sc, fgs_b200.RasterizerSettings(*st), to_chw=to_chw, clamp_output=clamp)
                b = ref.rasterize(*sc, ref.RasterizerSettings(*st), to_chw=to_chw, clamp_output=clamp)
                assert a.shape == b.shape and a.dtype == b.dtype
                l1 = (a - b).abs().mean().item()
                assert l1 <= IMAGE_L1, (n, to_chw, clamp, l1)
                if clamp:
                    assert round(O.psnr(a.cpu(), b.cpu()), 2) >= 60.0
                    assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0


def test_reference_python_bindings_run_over_this_C(cuda_device):
    """INTEGRATION.md option B on hardware: the reference's own, unmodified torch_bindings/*.py (the pip-installed copy
    under baseline/_ref) with `FasterGSCudaBackend._C` replaced by fgs_b200._C, against the all-reference stack."""
    ref_pkg = util.REF_DIR / 'FasterGSCudaBackend'
    if not (ref_pkg / 'torch_bindings' / 'rasterization.py').exists():
        pytest.skip('baseline/_ref not present')
    from fgs_b200 import _C, synthetic as syn
    import importlib

    def drop():
        for k in [k for k in sys.modules if k == 'FasterGSCudaBackend' or k.startswith('FasterGSCudaBackend.')]:
            del sys.modules[k]

    n, w, h = 20_000, 320, 240
    sc = syn.make_scene(n, 7, width=320).to('cuda')
    cam = syn.make_camera(7, w, h)
    st = util.settings_for(cam, bg=(0.3, 0.2, 0.1), device='cuda')
    gi = syn.make_grad_image(w, h, 7).cuda() / (3 * w * h)
    lrs = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]

    def train_two_steps(tb):
        params = [p.clone().requires_grad_(True) for p in sc]
        opt = tb.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs)], lr=0.0, eps=1e-15)
        dens = torch.zeros(2, n, device='cuda')
        imgs = []
        for _ in range(2):
            img = tb.diff_rasterize(*params, dens, tb.RasterizerSettings(*st))
            img.backward(gi)
            imgs.append(img.detach())
            grads = [p.grad.clone() for p in params]
            opt.step()
            opt.zero_grad()
        with torch.no_grad():
            hwc = tb.rasterize(*[p.detach() for p in params], tb.RasterizerSettings(*st), to_chw=False)
            scores = torch.zeros(n, device='cuda')
            tb.update_pruning_scores(scores, *[p.detach() for p in params], tb.RasterizerSettings(*st))
        torch.cuda.synchronize()
        return imgs, grads, [p.detach() for p in params], dens, hwc, scores

    drop()
    pkg = types.ModuleType('FasterGSCudaBackend')
    pkg.__path__ = [str(ref_pkg)]
    pkg._C = _C
    sys.modules['FasterGSCudaBackend'] = pkg
    sys.modules['FasterGSCudaBackend._C'] = _C
    try:
        tb = importlib.import_module('FasterGSCudaBackend.torch_bindings')
        assert str(ref_pkg) in tb.__file__
        assert sys.modules['FasterGSCudaBackend.torch_bindings.rasterization']._C is _C
        mine = train_two_steps(tb)
    finally:
        drop()
    ref = util.load_reference()  # the same Python files over the reference's own compiled _C
    theirs = train_two_steps(ref)
    drop()
    for a, b in zip(mine[0], theirs[0]):
        assert (a - b).abs().mean().item() <= IMAGE_L1
    for name, a, b in zip(NAMES, mine[1], theirs[1]):
        assert util.rel_err(a, b) <= GRAD_RTOL, (name, util.rel_err(a, b))
    for name, a, b, p0 in zip(NAMES, mine[2], theirs[2], sc):
        assert util.cos_sim(a - p0, b - p0) >= 0.999, name
    assert torch.equal(mine[3][0], theirs[3][0]) and util.rel_err(mine[3][1], theirs[3][1]) <= GRAD_RTOL
    assert (mine[4] - theirs[4]).abs().mean().item() <= IMAGE_L1
    assert util.rel_err(mine[5], theirs[5]) <= 2e-3


def test_alternating_resolutions_and_devices_keep_their_own_state(cuda_device):
    """Host state is per device and per shape class: a 1080p-like training view alternating with a larger evaluation view
    (and, with two GPUs, the same on a second device in the same process) neither re-runs the speculative tail after the
    first call of each shape nor changes any result."""
    import fgs_b200
    from fgs_b200 import _capi, synthetic as syn
    lib = _capi.load()
    devices = [torch.device('cuda', 0)] + ([torch.device('cuda', 1)] if torch.cuda.device_count() > 1 else [])
    shapes = [(30_000, 480, 270), (30_000, 960, 540)]
    first = {}
    reruns0 = {}
    for d in devices:
        with torch.cuda.device(d):
            reruns0[d] = lib.fgs_debug_get_forward_reruns()
    for rnd in range(4):
        for d in devices:
            for (n, w, h) in shapes:
                sc = syn.make_scene(n, 0, width=480).to(d)
                cam = syn.make_camera(1, w, h)
                st = util.settings_for(cam, device=d)
                gi = (syn.make_grad_image(w, h) / (3 * w * h)).to(d)
                img, grads = _run(fgs_b200, sc, st, gi)
                inf = fgs_b200.rasterize(*sc, fgs_b200.RasterizerSettings(*st), to_chw=True, clamp_output=False)
                torch.cuda.synchronize(d)
                assert (inf - img).abs().max().item() <= 1e-6
                key = (w, h)
                if key not in first:
                    first[key] = (img.cpu(), [g.cpu() for g in grads])
                else:
                    assert (img.cpu() - first[key][0]).abs().mean().item() <= 1e-6, (rnd, d, key)
                    for a, b in zip(grads, first[key][1]):
                        assert util.rel_err(a.cpu(), b) <= 1e-4
    for d in devices:
        with torch.cuda.device(d):
            # every (device, shape, training/inference) class starts cold exactly once; afterwards its own hint fits
            assert lib.fgs_debug_get_forward_reruns() == reruns0[d], 'a forward re-ran its speculative tail'


def test_deferred_sh_schedule_equals_fused_step(cuda_device):
    """diff_rasterize_fused(..., defer_sh=True): geometry pass without SH + colour kernel after the sorts + sh_rest Adam on
    the library's side stream.  Six training steps over different views give the images, parameters and moments of the
    plain fused path; inference renders, pruning scores and torch ops in between see a consistent sh_rest (joins)."""
    import fgs_b200
    from fgs_b200 import synthetic as syn
    n, w, h = 200_000, 960, 540
    sc = syn.make_scene(n, 3, width=960).to('cuda')
    lrs = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]
    cams = [syn.make_camera(k, w, h) for k in range(6)]
    gis = [(syn.make_grad_image(w, h, seed=k) / (3 * w * h)).cuda() for k in range(6)]

    def run(defer):
        params = [p.clone().requires_grad_(True) for p in sc]
        opt = fgs_b200.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs)], lr=0.0, eps=1e-15)
        imgs, extra = [], []
        for k in range(6):
            rs = fgs_b200.RasterizerSettings(*util.settings_for(cams[k], device='cuda'))
            img = fgs_b200.diff_rasterize_fused(*params, torch.empty(0), rs, opt, defer_sh=defer)
            img.backward(gis[k])
            opt.step()
            opt.zero_grad()
            imgs.append(img.detach())
            if k == 2:  # an evaluation render and pruning scores in the middle of training: both read sh_rest
                with torch.no_grad():
                    det = [p.detach() for p in params]
                    extra.append(fgs_b200.rasterize(*det, rs, to_chw=True))
                    scores = torch.zeros(n, device='cuda')
                    fgs_b200.update_pruning_scores(scores, *det, rs)
                    extra.append(scores)
            if k == 4:  # a plain torch read of sh_rest: the documented protocol is to join first
                fgs_b200.join_deferred()
                extra.append(params[5].detach().abs().sum().reshape(1))
        fgs_b200.join_deferred()
        torch.cuda.synchronize()
        return imgs, [p.detach() for p in params], [opt.state[p]['exp_avg'] for p in params], [opt.state[p]['exp_avg_sq'] for p in params], extra

    ia, pa, ma, va, xa = run(False)
    ib, pb, mb, vb, xb = run(True)
    for k, (a, b) in enumerate(zip(ia, ib)):
        assert (a - b).abs().mean().item() <= 1e-6, k
    for name, p0, a, b in zip(NAMES, sc, pa, pb):
        assert util.rel_err(b - p0, a - p0) <= 2e-3, (name, util.rel_err(b - p0, a - p0))  # Adam's first steps are sign-like in g
    for name, a, b in zip(NAMES, ma, mb):
        assert util.rel_err(b, a) <= 1e-4, (name, util.rel_err(b, a))
    for name, a, b in zip(NAMES, va, vb):
        assert util.rel_err(b, a) <= 1e-3, (name, util.rel_err(b, a))
    assert (xa[0] - xb[0]).abs().mean().item() <= 1e-6
    assert util.rel_err(xb[1], xa[1]) <= 1e-3
    assert util.rel_err(xb[2], xa[2]) <= 1e-5
