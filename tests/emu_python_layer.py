"""TEST INFRASTRUCTURE ONLY (run as a subprocess by tests/test_emulated_python_layer.py).

Runs the product's Python layer -- fgs_b200._C, torch_bindings (autograd Functions, FusedAdam), fgs_b200.loss,
fgs_b200.densification -- UNCHANGED on host tensors against the emulated build of the whole library
(FGS_B200_LIB -> tests/cuda_emu 'all'), with the three torch.cuda touch points of fgs_b200._capi replaced.  This checks the
argument marshalling, tensor shapes, autograd contracts and optimizer-state handling of the Python layer without a GPU;
the numerics are those of the emulated kernels (IEEE float32)."""
import contextlib
import math
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'faster-gaussian-splatting_b200')]

from tests import cuda_emu  # noqa: E402

os.environ['FGS_B200_LIB'] = str(cuda_emu.library_path('all'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from fgs_b200 import _capi  # noqa: E402

_capi.on_device = lambda t: True
_capi.current_stream = lambda: None
_capi.device_guard = lambda device: contextlib.nullcontext()

import fgs_b200  # noqa: E402
from fgs_b200 import densification as D  # noqa: E402
from fgs_b200 import loss as L  # noqa: E402
from fgs_b200 import synthetic as syn  # noqa: E402
from oracle import densify_oracle as DO  # noqa: E402
from oracle import loss_oracle as LO  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402
from tests import util  # noqa: E402

NAMES = ('means', 'scales', 'rotations', 'opacities', 'sh_coefficients_0', 'sh_coefficients_rest')
LRS = dict(means=1.6e-4, scales=5e-3, rotations=1e-3, opacities=2.5e-2, sh_coefficients_0=2.5e-3, sh_coefficients_rest=1.25e-4)


def check_rasterizer_api():
    n, w, h = 200, 48, 36
    sc, cam = syn.make_scene(n, 4), syn.make_camera(4, w, h)
    st = util.settings_for(cam, bg=(0.1, 0.2, 0.3))
    gi = syn.make_grad_image(w, h, 4)
    params = [p.clone().requires_grad_(True) for p in sc]
    dens = torch.zeros(2, n)
    rs = fgs_b200.RasterizerSettings(*st)
    img = fgs_b200.diff_rasterize(*params, dens, rs)
    assert tuple(img.shape) == (3, h, w) and not dens.requires_grad
    img.backward(gi)
    dens64 = torch.zeros(2, n, dtype=torch.float64)
    out, ograds = O.render_and_grads(list(sc), O.Settings(*st), gi, dtype=torch.float64, densification_info=dens64)
    assert (img.detach().double() - out.image).abs().mean().item() <= 2e-6
    for p, og in zip(params, ograds):
        assert tuple(p.grad.shape) == tuple(p.shape) and util.rel_err(p.grad, og) <= 2e-4
    assert torch.equal(dens[0].double(), dens64[0]) and util.rel_err(dens[1], dens64[1]) <= 2e-4
    # inference entry point: (H,W,3) clamped / (3,H,W)
    with torch.no_grad():
        hwc = fgs_b200.rasterize(*sc, rs, to_chw=False)
        chw = fgs_b200.rasterize(*sc, rs, to_chw=True, clamp_output=False)
    assert tuple(hwc.shape) == (h, w, 3) and (hwc.permute(2, 0, 1) - img.detach().clamp(0, 1)).abs().max().item() <= 1e-6
    assert (chw - img.detach()).abs().max().item() <= 1e-6
    # FusedAdam.step vs torch.optim.Adam on the same gradients; fused backward+Adam == backward-then-step
    opt = fgs_b200.FusedAdam([{'params': [p], 'lr': LRS[k], 'name': k} for p, k in zip(params, NAMES)], lr=0.0, eps=1e-15)
    ref_params = [p.detach().clone().requires_grad_(True) for p in params]
    ref = torch.optim.Adam([{'params': [p], 'lr': LRS[k]} for p, k in zip(ref_params, NAMES)], lr=0.0, eps=1e-15)
    for p, q in zip(params, ref_params):
        q.grad = p.grad.clone()
    opt.step()
    ref.step()
    for p, q, p0 in zip(params, ref_params, sc):
        assert util.rel_err(p.detach() - p0, q.detach() - p0) <= 2e-3
    opt.zero_grad()
    fparams = [p.clone().requires_grad_(True) for p in sc]
    fopt = fgs_b200.FusedAdam([{'params': [p], 'lr': LRS[k], 'name': k} for p, k in zip(fparams, NAMES)], lr=0.0, eps=1e-15)
    fimg = fgs_b200.diff_rasterize_fused(*fparams, torch.empty(0), rs, fopt)
    fimg.backward(gi)
    assert all(p.grad is None for p in fparams)
    fopt.step()  # no-op: no gradients
    for p, q, p0 in zip(fparams, params, sc):
        assert util.rel_err(p.detach() - p0, q.detach() - p0) <= 2e-3
        assert fopt.state[p]['step'] == 1 and util.rel_err(fopt.state[p]['exp_avg'], opt.state[q]['exp_avg']) <= 1e-4
    # pruning scores and the small operators go through the same marshalling
    scores = torch.zeros(n)
    fgs_b200.update_pruning_scores(scores, *sc, rs)
    assert util.rel_err(scores, O.pruning_scores(*sc, O.Settings(*st))) <= 2e-4
    # two forwards before their backwards (multi-view accumulation): the opaque per-call buffers must not interfere
    rs2 = fgs_b200.RasterizerSettings(*util.settings_for(syn.make_camera(1, w, h), bg=(0.1, 0.2, 0.3)))
    gi2 = syn.make_grad_image(w, h, 1)

    def accumulated(views):
        ps = [p.clone().requires_grad_(True) for p in sc]
        imgs = [fgs_b200.diff_rasterize(*ps, torch.empty(0), r) for r, _ in views]
        sum((im * g).sum() for im, (_, g) in zip(imgs, views)).backward()
        return [p.grad for p in ps]
    both, first, second = accumulated([(rs, gi), (rs2, gi2)]), accumulated([(rs, gi)]), accumulated([(rs2, gi2)])
    for x, y, z in zip(both, first, second):
        assert util.rel_err(x, y + z) <= 1e-5
    # the three small operators (3D filter, MCMC relocation, noise) against the compiled reference's golden and the oracle
    from tests import test_aux_ops as TA
    from fgs_b200 import _C
    d, filt = TA._inputs('cpu')
    got = TA._run(fgs_b200, d, filt, _C.add_noise)
    TA._compare(got, TA._oracle_run(TA._inputs('cpu')[0], filt), 'emulated kernels vs oracle')
    if TA.GOLDEN.exists():
        z = np.load(TA.GOLDEN)
        TA._compare(got, (z['filter_3d'], z['visibility'], z['new_opacities'], z['new_scales'], z['means']),
                    'emulated kernels vs reference golden')
    print('rasterizer API ok')


def check_loss_api():
    g = torch.Generator().manual_seed(0)
    w, h = 50, 37
    image = torch.rand(3, h, w, generator=g)
    target8 = (torch.rand(3, h, w, generator=g) * 255).round().to(torch.uint8)
    for tgt, l1w, dsw in ((target8, 0.8, 0.2), (target8.float() / 255, 1.0, 0.0)):
        terms, grad = L.l1_dssim_loss_and_grad(image, tgt, l1w, dsw)
        loss, l1, ds, g64 = LO.loss_and_grad(image, tgt, l1w, dsw)
        assert abs(float(terms[0]) - float(loss)) <= 2e-5 and util.rel_err(grad, g64) <= 2e-4
        x = image.clone().requires_grad_(True)
        out, terms2 = L.l1_dssim_loss(x, tgt, l1w, dsw, return_terms=True)
        assert out.dim() == 0 and not terms2.requires_grad
        (out * 0.37).backward()
        assert util.rel_err(x.grad, g64 * 0.37) <= 2e-4
    # a view at an odd offset is copied, not faulted on
    stack = torch.rand(2, 3, 5, 7, generator=g)
    t, gr = L.l1_dssim_loss_and_grad(stack[1], stack[0], 1.0, 0.0)
    assert abs(float(t[0]) - float((stack[1] - stack[0]).abs().mean())) <= 1e-6
    print('loss API ok')


def _optimizer(state):
    params = {k: torch.nn.Parameter(state[k].clone()) for k in NAMES}
    opt = fgs_b200.FusedAdam([{'params': [params[k]], 'lr': 1e-3, 'name': k} for k in NAMES], lr=0.0, eps=1e-15)
    for k in NAMES:
        opt.state[params[k]] = {'step': 7, 'exp_avg': state['exp_avg.' + k].clone(), 'exp_avg_sq': state['exp_avg_sq.' + k].clone()}
    return opt


def _state_of(opt, groups):
    out = {}
    for name, p in groups.items():
        out[name] = p.detach()
        st = opt.state[p]
        assert st['step'] == 7
        out['exp_avg.' + name], out['exp_avg_sq.' + name] = st['exp_avg'], st['exp_avg_sq']
    assert [g['params'][0] for g in opt.param_groups] == [groups[g['name']] for g in opt.param_groups] and len(opt.state) == 6
    return out


def check_densification_api():
    from tests.test_widen_densify import ADC, compare_states, fixed_randn, make_state
    n = 1300
    state, info = make_state(n, seed=8)
    want, z = DO.adaptive_density_control(state, info, ADC['grad_threshold'], ADC['min_opacity'], True, ADC['percent_dense'],
                                          ADC['extent'], fixed_randn(5))
    opt = _optimizer(state)
    groups = D.adaptive_density_control(opt, info, ADC['grad_threshold'], ADC['min_opacity'], True, ADC['percent_dense'],
                                        ADC['extent'], randn=lambda shape: z.reshape(shape))
    compare_states(_state_of(opt, groups), want, 'adc through the Python layer')
    g = torch.Generator().manual_seed(1)
    m = want['means'].shape[0]
    prune_mask = torch.rand(m, generator=g) < 0.25
    extra = torch.arange(2 * m, dtype=torch.float32).reshape(m, 2)
    groups, (extra_new,) = D.prune_param_groups(opt, ~prune_mask, extras=[extra])
    s1 = DO.prune(want, prune_mask)
    compare_states(_state_of(opt, groups), s1, 'prune')
    assert torch.equal(extra_new, extra[~prune_mask])
    means = groups['means'].detach()
    codes = D.morton_encode(means)
    assert torch.equal(codes, DO.morton_encode(means))
    order = D.morton_order(means)
    assert torch.equal(order, torch.sort(codes, stable=True).indices)
    groups = D.sort_param_groups(opt, order)
    s2 = DO.sort(s1, order)
    compare_states(_state_of(opt, groups), s2, 'morton sort')
    new = {k: torch.randn((33,) + tuple(state[k].shape[1:]), generator=g) for k in NAMES}
    groups = D.extend_param_groups(opt, new)
    s3 = DO.extend(s2, new)
    compare_states(_state_of(opt, groups), s3, 'extend')
    idx = torch.randperm(s3['means'].shape[0], generator=g)[:77]
    D.reset_state(opt, idx)
    compare_states(_state_of(opt, groups), DO.reset_state(s3, idx), 'reset_state')
    for p in groups.values():
        p.grad = torch.ones_like(p)
    opt.step()
    assert opt.state[groups['means']]['step'] == 8
    # 3DGS-MCMC relocation + growth (Model.py:368-457) with a shared deterministic sampler
    state2, _ = make_state(900, seed=21)
    state2['opacities'][::7] = -9.0  # dead Gaussians

    def sampler():
        gen = torch.Generator().manual_seed(3)
        return lambda weights, k: torch.multinomial(weights, k, replacement=True, generator=gen)
    want2 = DO.mcmc_densification(state2, 0.005, 930, sampler())
    opt2 = _optimizer(state2)
    groups2 = D.mcmc_densification(opt2, 0.005, 930, multinomial=sampler())
    got2 = _state_of(opt2, groups2)
    assert got2['means'].shape[0] == 930 == want2['means'].shape[0]
    for k in want2:
        if k in ('opacities', 'scales'):  # relocation_adjustment: alternating series, compared like tests/test_aux_ops.py
            rel = ((got2[k] - want2[k]).abs() / want2[k].abs().clamp_min(1e-12)).reshape(-1)
            assert float(rel.median()) <= 1e-5 and float((rel <= 5e-2).float().mean()) >= 0.97, k
        else:
            assert torch.equal(got2[k], want2[k]), k
    print('densification API ok')


def check_training_loop():
    """tests/test_widen_training.py's loop (render -> L1 + DSSIM -> backward -> step, densification, Morton reordering) at a
    size the emulator finishes in seconds: the plumbing between the modules, not the convergence."""
    from tests import test_widen_training as T
    T.N, T.W, T.H, T.VIEWS, T.ITERS = 150, 48, 36, 2, 6
    settings, targets, init, _ = T._problem('cpu')
    psnr0 = T._psnr(fgs_b200, init, settings, targets)
    params = T._train(fgs_b200, init, settings, targets, iters=6, loss_fn=lambda a, b: L.l1_dssim_loss(a, b, 0.8, 0.2),
                      densify_at=(3,), morton_at=(4,))
    p1 = T._psnr(fgs_b200, params, settings, targets)
    pf = T._psnr(fgs_b200, T._train(fgs_b200, init, settings, targets, iters=6, fused=True), settings, targets)
    pd = T._psnr(fgs_b200, T._train(fgs_b200, init, settings, targets, iters=6), settings, targets)
    assert math.isfinite(p1) and abs(pf - pd) <= 0.05, (psnr0, p1, pf, pd)
    print(f'training loop ok: PSNR {psnr0:.2f} -> {pd:.2f} (drop-in) / {pf:.2f} (fused) / {p1:.2f} (DSSIM + densify, N {T.N} -> {params["means"].shape[0]})')


def make_scene_4d(n, seed, width):
    sc = syn.make_scene(n, seed, width=width)
    g = torch.Generator().manual_seed(seed + 77)
    means4 = torch.cat([sc.means, torch.rand(n, 1, generator=g)], dim=1).contiguous()
    scales4 = torch.cat([sc.scales, torch.randn(n, 1, generator=g) * 0.3 - 1.2], dim=1).contiguous()
    rot_l = torch.randn(n, 4, generator=g)
    rot_r = torch.randn(n, 4, generator=g) * 0.3 + torch.tensor([1.0, 0.0, 0.0, 0.0])  # near the identity: mild space-time coupling
    return means4, scales4, rot_l.contiguous(), rot_r.contiguous(), sc.opacities, sc.sh_coefficients_0, sc.sh_coefficients_rest


def check_fourd_api():
    from fgs_b200 import fourd as F4
    from oracle import fourd_oracle as FO
    n, w, h, t = 250, 64, 48, 0.4
    params = make_scene_4d(n, 5, w)
    cam = syn.make_camera(5, w, h)
    st = util.settings_for(cam, bg=(0.05, 0.1, 0.15))
    gi = syn.make_grad_image(w, h, 5)
    leaves = [p.clone().requires_grad_(True) for p in params]
    dens = torch.zeros(2, n)
    img = F4.diff_rasterize_4d(*leaves, t, dens, fgs_b200.RasterizerSettings(*st))
    img.backward(gi)
    o_leaves = [p.double().clone().requires_grad_(True) for p in params]
    s = O.Settings(*st)
    s = s._replace(w2c=s.w2c.double(), cam_position=s.cam_position.double(), bg_color=s.bg_color.double())
    out = FO.render_4d(*o_leaves, t, s)
    out.image.backward(gi.double())
    vis = int(out.proj.visible.sum())
    assert 0 < vis < n and float(dens[0].sum()) == vis
    assert (img.detach().double() - out.image.detach()).abs().mean().item() <= 2e-6
    names = ('means4', 'scales4', 'rot_l', 'rot_r', 'opacities', 'sh0', 'sh_rest')
    for name, a, b in zip(names, leaves, o_leaves):
        assert a.grad is not None and util.rel_err(a.grad, b.grad) <= 5e-4, (name, util.rel_err(a.grad, b.grad))
    with torch.no_grad():
        inf = F4.rasterize_4d(*params, t, fgs_b200.RasterizerSettings(*st), to_chw=True, clamp_output=False)
    assert (inf - img.detach()).abs().max().item() <= 1e-6
    # another time slice renders a different image (the Gaussians move and fade)
    with torch.no_grad():
        other = F4.rasterize_4d(*params, 0.9, fgs_b200.RasterizerSettings(*st), to_chw=True, clamp_output=False)
    assert (other - inf).abs().mean().item() > 1e-4
    print(f'4D API ok: {vis} of {n} Gaussians visible at t = {t}')


def check_peer_adam_class():
    """fgs_b200.parallel.PeerAdam in a world of ONE rank (gloo), with torch's symmetric memory replaced by plain host
    tensors: the Python side of the class -- flat layout, parameter / gradient views, the rasterizer's backward writing
    its gradients straight into the symmetric gradient buffer (set_gradient_outputs), step / zero_grad / reset / close --
    against diff_rasterize + FusedAdam.step.  (The multi-rank slicing of the kernel is covered by
    tests/test_peer_adam.py::test_peer_adam_kernel_emulated_ranks.)"""
    import types
    import torch.distributed as dist
    from fgs_b200 import parallel
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29533')
    dist.init_process_group('gloo', rank=0, world_size=1)
    fake = types.ModuleType('torch.distributed._symmetric_memory')
    fake.empty = lambda n, dtype, device: torch.empty(n, dtype=dtype, device=device)

    class Handle:
        def __init__(self, t):
            self.buffer_ptrs = [t.data_ptr()]

        def barrier(self):
            pass
    fake.rendezvous = lambda t, group_name: Handle(t)
    sys.modules['torch.distributed._symmetric_memory'] = fake
    dist._symmetric_memory = fake
    n, w, h = 200, 48, 36
    sc, cam = syn.make_scene(n, 9), syn.make_camera(9, w, h)
    rs = fgs_b200.RasterizerSettings(*util.settings_for(cam))
    gis = [syn.make_grad_image(w, h, k) for k in range(2)]
    lrs = [LRS[k] for k in NAMES]
    pa = [p.clone().requires_grad_(True) for p in sc]
    oa = fgs_b200.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15)
    for it in range(2):
        fgs_b200.diff_rasterize(*pa, torch.empty(0), rs).backward(gis[it])
        oa.step()
        oa.zero_grad()
    ob = parallel.PeerAdam([p.clone() for p in sc], lrs)
    for it in range(2):
        fgs_b200.diff_rasterize(*ob.params, torch.empty(0), rs).backward(gis[it])
        assert all(p.grad.data_ptr() == g.data_ptr() for p, g in zip(ob.params, ob.grads)), 'gradient was copied, not adopted'
        ob.step()
        ob.zero_grad()
    ob.close()
    for a, b, p0 in zip(pa, ob.params, sc):
        assert util.rel_err(b.detach() - p0, a.detach() - p0) <= 2e-3
    ob.reset()
    assert ob.steps == 0 and float(ob.exp_avg.abs().sum()) == 0.0
    # after close() the backward allocates its own gradient tensors again
    q = [p.clone().requires_grad_(True) for p in sc]
    fgs_b200.diff_rasterize(*q, torch.empty(0), rs).backward(gis[0])
    assert q[0].grad.data_ptr() != ob.grads[0].data_ptr()
    dist.destroy_process_group()
    print('PeerAdam class ok')


def check_reference_bindings():
    """INTEGRATION.md option B, executed: the REFERENCE's own torch_bindings/*.py (autograd Function, FusedAdam, the small
    operators; read from /root/reference, which only exists in the development container) import `_C` from a package
    named FasterGSCudaBackend -- here that `_C` is this repo's fgs_b200._C over the emulated library.  Their unmodified
    Python then renders, back-propagates and steps; results against the oracle."""
    import types
    ref_pkg = Path('/root/reference/FasterGSCudaBackend/FasterGSCudaBackend')
    if not (ref_pkg / 'torch_bindings' / 'rasterization.py').exists():
        print('reference bindings: /root/reference not present, skipped')
        return
    from fgs_b200 import _C
    for k in [k for k in sys.modules if k == 'FasterGSCudaBackend' or k.startswith('FasterGSCudaBackend.')]:
        del sys.modules[k]
    pkg = types.ModuleType('FasterGSCudaBackend')
    pkg.__path__ = [str(ref_pkg)]
    pkg._C = _C
    sys.modules['FasterGSCudaBackend'] = pkg
    sys.modules['FasterGSCudaBackend._C'] = _C
    import importlib
    tb = importlib.import_module('FasterGSCudaBackend.torch_bindings')
    assert str(ref_pkg) in tb.__file__ and str(ref_pkg) in sys.modules['FasterGSCudaBackend.torch_bindings.rasterization'].__file__
    n, w, h = 200, 48, 36
    sc, cam = syn.make_scene(n, 7), syn.make_camera(7, w, h)
    st = util.settings_for(cam, bg=(0.3, 0.2, 0.1))
    gi = syn.make_grad_image(w, h, 7)
    params = [p.clone().requires_grad_(True) for p in sc]
    dens = torch.zeros(2, n)
    img = tb.diff_rasterize(*params, dens, tb.RasterizerSettings(*st))
    img.backward(gi)
    dens64 = torch.zeros(2, n, dtype=torch.float64)
    out, ograds = O.render_and_grads(list(sc), O.Settings(*st), gi, dtype=torch.float64, densification_info=dens64)
    assert (img.detach().double() - out.image).abs().mean().item() <= 2e-6
    for p, og in zip(params, ograds):
        assert util.rel_err(p.grad, og) <= 2e-4
    assert util.rel_err(dens[1], dens64[1]) <= 2e-4
    opt = tb.FusedAdam([{'params': [p], 'lr': LRS[k]} for p, k in zip(params, NAMES)], lr=0.0, eps=1e-15)
    ref_params = [p.detach().clone().requires_grad_(True) for p in params]
    ref_opt = torch.optim.Adam([{'params': [p], 'lr': LRS[k]} for p, k in zip(ref_params, NAMES)], lr=0.0, eps=1e-15)
    for p, q in zip(params, ref_params):
        q.grad = p.grad.clone()
    opt.step()
    ref_opt.step()
    for p, q, p0 in zip(params, ref_params, sc):
        assert util.rel_err(p.detach() - p0, q.detach() - p0) <= 2e-3
    with torch.no_grad():
        hwc = tb.rasterize(*sc, tb.RasterizerSettings(*st), to_chw=False)
    assert tuple(hwc.shape) == (h, w, 3) and (hwc.permute(2, 0, 1) - img.detach().clamp(0, 1)).abs().max().item() <= 1e-6
    scores = torch.zeros(n)
    tb.update_pruning_scores(scores, *sc, tb.RasterizerSettings(*st))
    assert util.rel_err(scores, O.pruning_scores(*sc, O.Settings(*st))) <= 2e-4
    # the small operators through the reference's own wrappers
    from tests import test_aux_ops as TA
    d, filt = TA._inputs('cpu')
    got = TA._run(tb, d, filt, _C.add_noise)
    TA._compare(got, TA._oracle_run(TA._inputs('cpu')[0], filt), "reference's bindings over this repo's _C vs oracle")
    for k in [k for k in sys.modules if k == 'FasterGSCudaBackend' or k.startswith('FasterGSCudaBackend.')]:
        del sys.modules[k]
    print("reference's own torch_bindings over fgs_b200._C ok")


def check_deferred_sh():
    """diff_rasterize_fused(..., defer_sh=True) through the Python layer: `_C.forward_ex` (deferred colour), `_C.backward_adam`
    with the sh_rest update deferred (on the emulator the side stream is synchronous), the keep-alive of the scratch the
    deferred kernel reads, `join_deferred`, and an inference render in between -- same parameters and moments as the plain
    fused path after four steps over two views."""
    from fgs_b200 import _C
    n, w, h = 180, 48, 36
    sc = syn.make_scene(n, 9)
    cams = [syn.make_camera(k, w, h) for k in range(2)]
    rs = [fgs_b200.RasterizerSettings(*util.settings_for(c, bg=(0.2, 0.1, 0.3))) for c in cams]
    gis = [syn.make_grad_image(w, h, k) for k in range(2)]

    def run(defer):
        params = [p.clone().requires_grad_(True) for p in sc]
        opt = fgs_b200.FusedAdam([{'params': [p], 'lr': LRS[k]} for p, k in zip(params, NAMES)], lr=0.0, eps=1e-15)
        imgs = []
        for it in range(4):
            img = fgs_b200.diff_rasterize_fused(*params, torch.empty(0), rs[it % 2], opt, defer_sh=defer)
            img.backward(gis[it % 2])
            opt.step()
            opt.zero_grad()
            imgs.append(img.detach().clone())
            if it == 1:
                with torch.no_grad():
                    imgs.append(fgs_b200.rasterize(*[p.detach() for p in params], rs[0], to_chw=True))
        fgs_b200.join_deferred('cpu')
        return imgs, [p.detach().clone() for p in params], [opt.state[p]['exp_avg'].clone() for p in params]

    ia, pa, ma = run(False)
    ib, pb, mb = run(True)
    assert (not _C._deferred_keepalive) is False  # the deferred path kept its scratch alive
    for a, b in zip(ia, ib):
        assert (a - b).abs().max().item() <= 1e-6
    for k, a, b, p0 in zip(NAMES, pa, pb, sc):
        assert util.rel_err(b - p0, a - p0) <= 2e-3, k
    for k, a, b in zip(NAMES, ma, mb):
        assert util.rel_err(b, a) <= 1e-4, k
    print('deferred SH schedule through the Python layer ok')


def check_gpu_test_code():
    """The bodies of the `-m gpu` tests of the widened rows, executed here with device 'cpu' (and smaller sizes where the
    test takes them as parameters): validates the TEST code that will first meet a GPU at the end of the round."""
    from tests import test_widen_densify as TD
    from tests import test_widen_fourd as TF
    from tests import test_widen_loss as TL
    from tests import test_widen_overlap_gpu as TO
    dev = torch.device('cpu')
    TL.test_loss_gpu_vs_oracle(dev, 70, 45, False, 0.8, 0.2)
    TL.test_loss_gpu_vs_oracle(dev, 16, 12, True, 1.0, 0.0)
    TD.test_adc_gpu_vs_oracle(dev, 2500, True)
    TD.test_prune_sort_extend_reset_morton_gpu(dev, n=4000)
    TO.test_overlapped_adam_schedule_matches_default(dev, 300, 48, 36)
    TF.test_fourd_forward_backward_vs_oracle(dev)
    print('GPU test code ok')


if __name__ == '__main__':
    torch.cuda.synchronize = lambda *a, **k: None  # the training helper ends with a synchronize
    which = sys.argv[1:] or ['rasterizer', 'loss', 'densification', 'fourd', 'training']
    for name in which:
        {'rasterizer': check_rasterizer_api, 'loss': check_loss_api, 'densification': check_densification_api,
         'fourd': check_fourd_api, 'training': check_training_loop, 'gpu_tests': check_gpu_test_code, 'peer_adam': check_peer_adam_class,
         'reference_bindings': check_reference_bindings, 'deferred_sh': check_deferred_sh}[name]()
    print('ALL OK')
