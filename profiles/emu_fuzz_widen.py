"""Randomised sweep of the loss and densification kernels on the CPU emulator against their oracles.

    python profiles/emu_fuzz_widen.py <seed> <seconds>
"""
import math
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'faster-gaussian-splatting_b200')]
import numpy as np
import torch

from oracle import densify_oracle as DO
from tests import test_widen_densify as TD
from tests import test_widen_loss as TL

rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 300)
loss_lib = TL._emu()
dens_lib, RowTensor = TD._emu()
n_loss = n_dens = bad = 0
while time.time() < t_end:
    # ---- loss ----
    w, h = rng.randint(1, 90), rng.randint(1, 70)
    u8 = rng.random() < 0.5
    l1w, dsw = rng.choice([(0.8, 0.2), (1.0, 0.0), (0.0, 1.0), (0.3, 0.7)])
    image, target = TL._images(w, h, seed=rng.randrange(10000), u8=u8)
    img, tgt = image.numpy(), target.numpy()
    scratch = np.zeros(loss_lib.fgs_loss_scratch_bytes(w, h) + 128, dtype=np.uint8)
    sp = scratch.ctypes.data + (-scratch.ctypes.data) % 128
    loss3, grad = np.full(3, np.nan, np.float32), np.full((3, h, w), np.nan, np.float32)
    assert loss_lib.fgs_loss_forward_backward(img.ctypes.data, tgt.ctypes.data, int(u8), w, h, l1w, dsw, sp, loss3.ctypes.data,
                                              grad.ctypes.data, None) == 0
    try:
        TL._check(loss3, grad, image, target, l1w, dsw)
        n_loss += 1
    except AssertionError as e:
        bad += 1
        print('BAD loss', w, h, u8, l1w, dsw, e, flush=True)
    # ---- adaptive density control ----
    n = rng.choice([1, 31, 32, 33, 1023, 1024, 1025, 2049, 3000])
    prune_large = rng.random() < 0.5
    state, info = TD.make_state(n, seed=rng.randrange(10000))
    thr = rng.choice([0.0002, 0.00005, 0.001])
    want, z = DO.adaptive_density_control(state, info, thr, 0.005, prune_large, 0.01, 4.0, TD.fixed_randn(5))
    flags = np.zeros(n, dtype=np.uint8)
    assert dens_lib.fgs_adc_flags(info.numpy().ctypes.data, state['scales'].numpy().ctypes.data, state['rotations'].numpy().ctypes.data,
                                  state['opacities'].numpy().ctypes.data, n, thr, math.log(0.04), math.log(0.005 / 0.995),
                                  int(prune_large), math.log(0.4), flags.ctypes.data, None) == 0
    plan, totals = TD.emu_plan(dens_lib, flags)
    zs = np.ascontiguousarray(z.numpy().astype(np.float32))
    try:
        assert 2 * totals[3] == z.shape[0]
        if plan.shape[0] > 0:
            got = TD.emu_move_rows(dens_lib, RowTensor, state, plan, plan.shape[0],
                                   split=(state['scales'].numpy(), state['rotations'].numpy(), zs))
            TD.compare_states(got, want)
        else:
            assert want['means'].shape[0] == 0
        n_dens += 1
    except AssertionError as e:
        bad += 1
        print('BAD adc', n, prune_large, thr, e, flush=True)
print('loss cases', n_loss, 'adc cases', n_dens, 'bad', bad)
