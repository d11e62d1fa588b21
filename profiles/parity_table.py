#!/usr/bin/env python
"""Parity noise floor on hardware: same inputs through
  (i)   the compiled reference twice            -> run-to-run spread of the reference (float atomics)
  (ii)  the compiled reference vs the fp64 oracle -> what the reference's own fast-math costs
  (iii) this repo vs the fp64 oracle
  (iv)  this repo vs the compiled reference
  (v)   this repo twice
and prints one markdown table per case: image mean-L1 and, per gradient tensor, |a-b|_1/|b|_1.

Cases: C1 (1 k Gaussians, 128x128; three seeds, with/without anti-aliasing; oracle = whole image, fp64) and C2
(1 M Gaussians, 1920x1080; the fp64 oracle is dense, so at C2 it is evaluated for a loss restricted to a pixel window:
grad_image is zero outside the window, which makes the six gradients a function of the window's pixels only).

    python profiles/parity_table.py > gpurun_out/r2_parity.md      (GPU box; baseline/_ref must have travelled)

Test infrastructure: imports oracle/ and baseline/_ref as checkers only."""
import json
import math
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'faster-gaussian-splatting_b200')]

NAMES = ['means', 'scales', 'rotations', 'opacities', 'sh0', 'sh_rest']


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).abs().sum() / b.abs().sum().clamp_min(1e-300))


def run(be, sc, st, gi):
    params = [p.clone().requires_grad_(True) for p in sc]
    img = be.diff_rasterize(*params, torch.empty(0), be.RasterizerSettings(*st))
    img.backward(gi)
    torch.cuda.synchronize()
    return img.detach(), [p.grad.detach() for p in params]


def oracle(sc, st, gi, window=None):
    from oracle import torch_oracle as O
    cpu = [p.detach().cpu() for p in sc]
    s = O.Settings(*[t.cpu() if torch.is_tensor(t) else t for t in st])
    if window is None:
        out, grads = O.render_and_grads(cpu, s, gi.cpu(), dtype=torch.float64)
        return out.image, grads
    dt = torch.float64
    leaves = [p.to(dt).clone().requires_grad_(True) for p in cpu]
    s = s._replace(w2c=s.w2c.to(dt), cam_position=s.cam_position.to(dt), bg_color=s.bg_color.to(dt))
    out = O.render(*leaves, s, window=window, pixel_chunk=2048)
    x0, y0, x1, y1 = window
    out.image.backward(gi.cpu().to(dt)[:, y0:y1, x0:x1])
    grads = [torch.zeros_like(l) if l.grad is None else l.grad for l in leaves]
    inv = ~out.proj.visible
    grads = [g.masked_fill(inv.reshape(-1, *([1] * (g.dim() - 1))), 0.0) for g in grads]
    return out.image.detach(), grads


def table(title, cols, rows, notes=()):
    print(f'\n### {title}\n')
    print('| quantity | ' + ' | '.join(cols) + ' |')
    print('|---|' + '---|' * len(cols))
    for name, vals in rows:
        print(f'| {name} | ' + ' | '.join('--' if v is None else f'{v:.2e}' for v in vals) + ' |')
    for n in notes:
        print(f'\n{n}')
    sys.stdout.flush()


def case(ref, mine, sc, st, gi, title, window=None, with_oracle=True):
    from oracle import torch_oracle as O
    r1 = run(ref, sc, st, gi) if ref is not None else None
    r2 = run(ref, sc, st, gi) if ref is not None else None
    m1 = run(mine, sc, st, gi)
    m2 = run(mine, sc, st, gi)
    o = None
    t0 = time.time()
    if with_oracle:
        o = oracle(sc, st, gi, window)
    t_or = time.time() - t0
    cols = ['ref vs ref (i)', 'ref vs fp64 oracle (ii)', 'mine vs fp64 oracle (iii)', 'mine vs ref (iv)', 'mine vs mine (v)']

    def img_l1(a, b, win=None):
        if a is None or b is None:
            return None
        a, b = a.double().cpu(), b.double().cpu()
        return float((a - b).abs().mean())

    def crop(img):
        if window is None or img is None:
            return img
        x0, y0, x1, y1 = window
        return img[:, y0:y1, x0:x1]

    rows = []
    rows.append(('image mean-L1' + (' (window)' if window else ''), [
        img_l1(r1[0], r2[0]) if r1 else None,
        img_l1(crop(r1[0]), o[0]) if (r1 and o) else None,
        img_l1(crop(m1[0]), o[0]) if o else None,
        img_l1(m1[0], r1[0]) if r1 else None,
        img_l1(m1[0], m2[0])]))
    worst = {}
    for k, name in enumerate(NAMES):
        vals = [rel(r1[1][k], r2[1][k]) if r1 else None,
                rel(r1[1][k], o[1][k]) if (r1 and o) else None,
                rel(m1[1][k], o[1][k]) if o else None,
                rel(m1[1][k], r1[1][k]) if r1 else None,
                rel(m1[1][k], m2[1][k])]
        rows.append((f'grad {name} rel-L1', vals))
        worst[name] = vals
    notes = []
    if r1:
        pm = O.psnr(m1[0].clamp(0, 1).cpu(), r1[0].clamp(0, 1).cpu())
        notes.append(f'PSNR(mine, ref) = {pm:.2f} dB; image max |diff| = {float((m1[0] - r1[0]).abs().max()):.2e}')
    if with_oracle:
        notes.append(f'oracle time {t_or:.1f} s' + (f', window {window}' if window else ''))
    table(title, cols, rows, notes)
    return worst


def main():
    import fgs_b200
    from fgs_b200 import synthetic as syn
    from tests import util
    assert torch.cuda.is_available()
    torch.set_num_threads(min(32, torch.get_num_threads()))
    ref = util.load_reference()
    print('# Parity noise floor, measured on', torch.cuda.get_device_name(0))
    print('\nreference =', 'baseline/_ref (unmodified, sm_100a build, -use_fast_math)' if ref is not None else 'NOT AVAILABLE')
    print('\nAll gradient numbers are |a-b|_1 / |b|_1 per tensor; image numbers are mean |a-b| over 3xHxW.')
    summary = {}
    # ---- C1 -------------------------------------------------------------------------------------------------------
    for seed in (0, 1, 2):
        for aa in (False, True):
            sc = syn.make_scene(1000, seed).to('cuda')
            cam = syn.make_camera(seed, 128, 128)
            st = util.settings_for(cam, bg=(0.2, 0.3, 0.1), aa=aa, device='cuda')
            gi = syn.make_grad_image(128, 128, seed).cuda()
            summary[f'C1 seed {seed} aa {aa}'] = case(ref, fgs_b200, sc, st, gi, f'C1: 1 k Gaussians, 128x128, seed {seed}, anti-aliasing {aa}')
    # ---- mid size: 50 k Gaussians at 480x270 (same footprint statistics as C2), full-image oracle ---------------------
    sc = syn.make_scene(50_000, 0, width=480).to('cuda')
    cam = syn.make_camera(0, 480, 270)
    st = util.settings_for(cam, device='cuda')
    gi = syn.make_grad_image(480, 270).cuda() / (3 * 480 * 270)
    win = (192, 96, 288, 168)
    giw = torch.zeros_like(gi)
    giw[:, win[1]:win[3], win[0]:win[2]] = gi[:, win[1]:win[3], win[0]:win[2]]
    summary['mid'] = case(ref, fgs_b200, sc, st, giw, 'mid: 50 k Gaussians, 480x270, loss restricted to a 96x72 window (fp64 oracle on that window)', window=win)
    # ---- C2 -------------------------------------------------------------------------------------------------------
    n, w, h = 1_000_000, 1920, 1080
    sc = syn.make_scene(n, 0).to('cuda')
    cam = syn.make_camera(0, w, h)
    st = util.settings_for(cam, device='cuda')
    gi = syn.make_grad_image(w, h).cuda() / (3 * w * h)
    summary['C2 full'] = case(ref, fgs_b200, sc, st, gi, 'C2: 1 M Gaussians, 1920x1080, dL/dimage ~ N(0,1)/(3HW) on the whole image (no oracle: dense fp64 is out of reach)',
                              with_oracle=False)
    for win in ((912, 504, 1008, 576), (300, 200, 396, 272)):
        giw = torch.zeros_like(gi)
        x0, y0, x1, y1 = win
        giw[:, y0:y1, x0:x1] = gi[:, y0:y1, x0:x1]
        summary[f'C2 window {win}'] = case(ref, fgs_b200, sc, st, giw, f'C2 scene, loss restricted to the {x1 - x0}x{y1 - y0} window at ({x0},{y0}) (fp64 oracle on that window)', window=win)
    # the L1 training loss of bench.py: dL/dimage = sign(img - gt)/(3HW)
    gt = syn.make_target(w, h).cuda()
    img, _ = run(fgs_b200, sc, st, gi)
    gl1 = torch.sign(img - gt) / (3 * w * h)
    summary['C2 L1'] = case(ref, fgs_b200, sc, st, gl1, 'C2 scene, dL/dimage = sign(img - target)/(3HW) (the benchmark loss)', with_oracle=False)
    print('\n<!-- json ' + json.dumps(summary) + ' -->')


if __name__ == '__main__':
    main()
