"""Fused training loss (L1 + DSSIM; reference Loss.py:15-16,26-33).

CPU: the oracle against an independent dense evaluation; the kernel SOURCE of csrc/loss.cu executed by the test-only
CUDA emulator (tests/cuda_emu) against the oracle -- forward, backward with a device-side dL/dloss, one-call
forward+backward, float and uint8 targets, ragged image sizes.  GPU: the same comparisons through the C ABI / the
Python operator on the real kernels."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import loss_oracle as LO


def _images(w, h, seed=0, u8=False):
    g = torch.Generator().manual_seed(seed)
    # a smooth image plus noise, and a target correlated with it: SSIM well inside (0, 1)
    yy, xx = torch.meshgrid(torch.linspace(0, 3, h), torch.linspace(0, 4, w), indexing='ij')
    base = torch.stack([0.5 + 0.4 * torch.sin(xx + c) * torch.cos(yy - c) for c in range(3)])
    image = (base + 0.08 * torch.randn(3, h, w, generator=g)).float().contiguous()
    target = (base + 0.05 * torch.randn(3, h, w, generator=g)).clamp(0, 1).float().contiguous()
    if u8:
        target = (target * 255).round().to(torch.uint8).contiguous()
    return image, target


def test_loss_oracle_selfcheck():
    image, target = _images(37, 29)
    a = LO.ssim_map(image.double(), target.double())
    b = LO.ssim_map_dense(image.double(), target.double())
    assert torch.allclose(a, b, rtol=1e-10, atol=1e-12)
    assert float(LO.ssim_map(image.double(), image.double()).min()) > 1.0 - 1e-9  # SSIM(x, x) == 1
    loss, l1, ds, g = LO.loss_and_grad(image, target, 1.0, 0.0)
    assert torch.allclose(g, torch.sign(image - target).double() / image.numel())
    assert abs(float(loss) - float((image - target).abs().mean())) < 1e-7
    # uint8 targets are value / 255
    _, t8 = _images(37, 29, u8=True)
    l1_8, _ = LO.loss_terms(image, t8)
    assert abs(float(l1_8) - float((image - t8.float() / 255).abs().mean())) < 1e-7


def _emu():
    from tests import cuda_emu
    lib = cuda_emu.build('loss')
    lib.fgs_loss_scratch_bytes.restype = C.c_size_t
    lib.fgs_loss_scratch_bytes.argtypes = [C.c_int32, C.c_int32]
    P = C.c_void_p
    lib.fgs_loss_forward.argtypes = [P, P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, P, P, P]
    lib.fgs_loss_backward.argtypes = [P, P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, P, P, P, P]
    lib.fgs_loss_forward_backward.argtypes = [P, P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, P, P, P, P]
    return lib


def _check(loss3, grad, image, target, l1w, dsw, grad_loss=1.0, what=''):
    loss, l1, ds, g = LO.loss_and_grad(image, target, l1w, dsw, grad_loss)
    assert abs(float(loss3[1]) - float(l1)) <= 1e-6 + 1e-5 * abs(float(l1)), what
    if dsw != 0.0:
        assert abs(float(loss3[2]) - float(ds)) <= 2e-5, (what, float(loss3[2]), float(ds))
    assert abs(float(loss3[0]) - float(loss)) <= 2e-5, what
    g = g.float()
    err = float((torch.as_tensor(grad).double() - g.double()).abs().sum() / g.double().abs().sum())
    assert err <= 2e-4, (what, err)


@pytest.mark.parametrize('w,h,u8,l1w,dsw', [
    (70, 45, False, 0.8, 0.2),   # ragged in both directions, three tile columns
    (64, 32, True, 0.8, 0.2),    # exact tiles, uint8 target
    (33, 7, False, 0.0, 1.0),    # image smaller than the window in one direction
    (50, 21, False, 1.0, 0.0),   # L1 only (streaming kernel), 3*50*21 = 3150 values: not a multiple of 4
    (16, 12, True, 1.0, 0.0),
])
def test_loss_kernels_emulated(w, h, u8, l1w, dsw):
    lib = _emu()
    image, target = _images(w, h, seed=w + h, u8=u8)
    img = image.numpy()
    tgt = target.numpy()
    scratch = np.zeros(lib.fgs_loss_scratch_bytes(w, h) + 128, dtype=np.uint8)
    sp = scratch.ctypes.data + (-scratch.ctypes.data) % 128
    scratch[:] = 0xAB  # the entry points must not rely on zeroed scratch
    loss3 = np.full(3, np.nan, dtype=np.float32)
    grad = np.full((3, h, w), np.nan, dtype=np.float32)
    rc = lib.fgs_loss_forward_backward(img.ctypes.data, tgt.ctypes.data, int(u8), w, h, l1w, dsw, sp, loss3.ctypes.data,
                                       grad.ctypes.data, None)
    assert rc == 0
    _check(loss3, grad, image, target, l1w, dsw, what='forward_backward')
    # two-call form with dL/dloss read from "device" memory
    loss3b = np.full(3, np.nan, dtype=np.float32)
    gradb = np.full((3, h, w), np.nan, dtype=np.float32)
    gl = np.array([0.37], dtype=np.float32)
    assert lib.fgs_loss_forward(img.ctypes.data, tgt.ctypes.data, int(u8), w, h, l1w, dsw, sp, loss3b.ctypes.data, None) == 0
    assert lib.fgs_loss_backward(img.ctypes.data, tgt.ctypes.data, int(u8), w, h, l1w, dsw, sp, gl.ctypes.data,
                                 gradb.ctypes.data, None) == 0
    assert np.array_equal(loss3, loss3b)
    _check(loss3b, gradb, image, target, l1w, dsw, grad_loss=0.37, what='forward + backward')
    assert np.isfinite(grad).all() and np.isfinite(gradb).all()  # every pixel written


def test_loss_emulated_argument_errors():
    lib = _emu()
    lib.fgs_last_error.restype = C.c_char_p
    a = np.zeros(16, dtype=np.float32)
    assert lib.fgs_loss_forward(None, a.ctypes.data, 0, 2, 2, 1.0, 0.0, a.ctypes.data, a.ctypes.data, None) == -1
    assert b'NULL' in lib.fgs_last_error()
    assert lib.fgs_loss_forward(a.ctypes.data, a.ctypes.data, 0, 0, 2, 1.0, 0.0, a.ctypes.data, a.ctypes.data, None) == -1


# ---- GPU: the real kernels ----------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('w,h,u8,l1w,dsw', [
    (70, 45, False, 0.8, 0.2), (64, 32, True, 0.8, 0.2), (33, 7, False, 0.0, 1.0), (50, 21, False, 1.0, 0.0),
    (16, 12, True, 1.0, 0.0), (517, 389, False, 0.8, 0.2),
])
def test_loss_gpu_vs_oracle(cuda_device, w, h, u8, l1w, dsw):
    from fgs_b200 import loss as L
    image, target = _images(w, h, seed=w + h, u8=u8)
    xi, ti = image.to(cuda_device), target.to(cuda_device)
    terms, grad = L.l1_dssim_loss_and_grad(xi, ti, l1w, dsw)
    _check(terms.cpu(), grad.cpu(), image, target, l1w, dsw, what='loss_and_grad')
    # autograd form, scaled upstream gradient
    x = xi.clone().requires_grad_(True)
    loss = L.l1_dssim_loss(x, ti, l1w, dsw)
    (loss * 0.37).backward()
    ref = LO.loss_and_grad(image, target, l1w, dsw, 0.37)
    assert abs(float(loss) - float(ref[0])) <= 2e-5
    err = float((x.grad.cpu().double() - ref[3]).abs().sum() / ref[3].abs().sum())
    assert err <= 2e-4, err


@pytest.mark.gpu
def test_loss_gpu_full_size_properties(cuda_device):
    """BASELINE.json's 1080p: size-independent properties instead of the (slow) oracle."""
    from fgs_b200 import loss as L
    w, h = 1920, 1080
    g = torch.Generator(device='cpu').manual_seed(3)
    image = torch.rand(3, h, w, generator=g).to(cuda_device)
    target = (torch.rand(3, h, w, generator=g) * 255).round().to(torch.uint8).to(cuda_device)
    tf = target.float().div(255.0)
    terms, grad = L.l1_dssim_loss_and_grad(image, target, 0.8, 0.2)
    terms_f, grad_f = L.l1_dssim_loss_and_grad(image, tf, 0.8, 0.2)
    # uint8 target == the same target converted by torch
    assert float((grad - grad_f).abs().sum() / grad_f.abs().sum()) <= 1e-5 and torch.allclose(terms, terms_f, rtol=0, atol=1e-6)
    assert abs(float(terms[1]) - float((image - tf).abs().mean())) <= 1e-6
    # SSIM(x, x) == 1 -> DSSIM 0 and zero DSSIM gradient (up to fp32 cancellation in E[x^2] - mu^2)
    t2, g2 = L.l1_dssim_loss_and_grad(tf, tf, 0.0, 1.0)
    assert abs(float(t2[2])) <= 5e-6 and float(g2.abs().max()) <= 1e-9
    # directional derivative along sign(grad) == sum |grad| (central difference of the fp32 loss; signal ~3, noise ~1e-4)
    _, gd = L.l1_dssim_loss_and_grad(image, target, 0.0, 1.0)
    d = torch.sign(gd)
    eps = 2e-3
    lp, _ = L.l1_dssim_loss_and_grad(image + eps * d, target, 0.0, 1.0)
    lm, _ = L.l1_dssim_loss_and_grad(image - eps * d, target, 0.0, 1.0)
    fd = (float(lp[0]) - float(lm[0])) / (2 * eps)
    an = float(gd.double().abs().sum())
    assert abs(fd - an) <= 1e-2 * an, (fd, an)
