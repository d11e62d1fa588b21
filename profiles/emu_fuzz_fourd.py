"""Randomised sweep of the 4D path on the CPU emulator (tests/cuda_emu): fgs_slice4d_forward -> fgs_forward / fgs_backward in
covariance-input mode -> fgs_slice4d_backward, against oracle/fourd_oracle.py (torch autograd, float64).

    python profiles/emu_fuzz_fourd.py <seed> <seconds>
"""
import ctypes as C
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'faster-gaussian-splatting_b200')]
import numpy as np
import torch

from fgs_b200 import synthetic as syn
from oracle import fourd_oracle as FO
from oracle import torch_oracle as O
from tests import cuda_emu, emu_driver as E, util

lib = E.load()
sl = cuda_emu.build('slice4d')
P = C.c_void_p
sl.fgs_slice4d_forward.argtypes = [P, P, P, P, P, C.c_float, C.c_int32, P, P, P, P]
sl.fgs_slice4d_backward.argtypes = [P, P, P, P, P, C.c_float, C.c_int32, P, P, P, P, P, P, P, P, P]
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 300)
it = bad = 0
while time.time() < t_end:
    it += 1
    n = rng.choice([1, 5, 33, 100, 257])
    w, h = rng.choice([(16, 12), (33, 25), (64, 48), (96, 30)])
    seed = rng.randrange(1000)
    t = rng.random()
    coupling = rng.choice([0.0, 0.3, 1.0])  # 1.0: fully random right quaternion (strong space-time mixing)
    bases = rng.choice([1, 4, 16])
    sc = syn.make_scene(n, seed, width=w)
    g = torch.Generator().manual_seed(seed + 1)
    means4 = torch.cat([sc.means, torch.rand(n, 1, generator=g)], 1).contiguous()
    scales4 = torch.cat([sc.scales, torch.randn(n, 1, generator=g) * 0.4 - 1.0], 1).contiguous()
    rot_l = torch.randn(n, 4, generator=g).contiguous()
    rot_r = (torch.randn(n, 4, generator=g) * coupling + torch.tensor([1.0, 0, 0, 0]) * (1 - coupling)).contiguous()
    params = [means4, scales4, rot_l, rot_r, sc.opacities, sc.sh_coefficients_0, sc.sh_coefficients_rest]
    cam = syn.make_camera(seed, w, h)
    st = util.settings_for(cam, active_sh_bases=bases, bg=(rng.random(), rng.random(), rng.random()))
    gi = syn.make_grad_image(w, h, seed)
    tag = f'n={n} {w}x{h} seed={seed} t={t:.3f} coupling={coupling} bases={bases}'
    ins = [E._copy_aligned(x.numpy()) for x in params[:5]]
    m3, c6, lg = (E._copy_aligned(np.zeros((n, k), np.float32)) for k in (3, 6, 1))
    assert sl.fgs_slice4d_forward(*[a.ctypes.data for a in ins], t, n, m3.ctypes.data, c6.ctypes.data, lg.ctypes.data, None) == 0
    p3 = [torch.from_numpy(np.array(m3)), torch.from_numpy(np.array(c6)), None, torch.from_numpy(np.array(lg)), params[5], params[6]]
    img, state, scr, inp = E.forward(lib, p3, st)
    grads3 = E.backward(lib, img, gi, state, scr, inp)
    gin = [E._copy_aligned(grads3[k].numpy()) for k in (0, 1, 3)]
    gout = [E._copy_aligned(np.full(x.shape, np.nan, np.float32)) for x in params[:5]]
    assert sl.fgs_slice4d_backward(*[a.ctypes.data for a in ins], t, n, *[a.ctypes.data for a in gin], *[a.ctypes.data for a in gout], None) == 0
    leaves = [x.double().clone().requires_grad_(True) for x in params]
    s = O.Settings(*st)
    s = s._replace(w2c=s.w2c.double(), cam_position=s.cam_position.double(), bg_color=s.bg_color.double())
    out = FO.render_4d(*leaves, t, s)
    out.image.backward(gi.double())
    l1 = float((torch.from_numpy(np.array(img)).double() - out.image.detach()).abs().mean())
    got = [torch.from_numpy(np.array(a)) for a in gout] + [grads3[4], grads3[5]]
    errs = []
    for a, b in zip(got, leaves):
        ref = torch.zeros_like(b) if b.grad is None else b.grad
        errs.append(util.rel_err(a, ref) if float(ref.abs().sum()) > 0 else float(a.abs().sum()))
    ok = np.isfinite(np.array(img)).all() and all(torch.isfinite(a).all() for a in got) and l1 <= 2e-5 and max(errs) <= 5e-3 \
        and state.n_visible == int(out.proj.visible.sum())
    bad += 0 if ok else 1
    print(('ok  ' if ok else 'BAD ') + tag + f' V={state.n_visible} l1={l1:.2e} maxerr={max(errs):.2e}', flush=True)
print('iterations', it, 'bad', bad)
