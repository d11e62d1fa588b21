"""TEST INFRASTRUCTURE ONLY: drives the emulated build of the whole rasterizer (tests/cuda_emu, `build('full')`) through
the same C ABI as the product (include/fgs_b200.h), with numpy arrays standing in for device memory."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from fgs_b200 import _capi


def load(sanitize_threads: bool = False):
    from tests import cuda_emu
    lib = cuda_emu.build('full', sanitize_threads=sanitize_threads)
    for name, (res, args) in _capi.SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


def _aligned(nbytes: int, align: int = 256) -> np.ndarray:
    raw = np.zeros(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


class Scratch:
    """The three caller-owned growable buffers of fgs_forward (resize callbacks)."""

    def __init__(self):
        self.buf = {}
        self.cb = _capi.ResizeFn(self._resize)

    def _resize(self, user, nbytes):
        user = int(user)
        if int(nbytes) == 0 and user in self.buf:
            return self.buf[user].ctypes.data
        self.buf[user] = _aligned(max(int(nbytes), 1))
        self.buf[user][:] = 0xCD  # the library must not rely on zeroed scratch
        return self.buf[user].ctypes.data


def _f32(t) -> np.ndarray:
    return np.ascontiguousarray(torch.as_tensor(t).detach().cpu().numpy().astype(np.float32))


def _copy_aligned(a: np.ndarray) -> np.ndarray:
    out = _aligned(max(a.nbytes, 16)).view(np.float32)[:a.size].reshape(a.shape)
    out[...] = a
    return out


def settings_struct(st, total_rest):
    return _capi.Settings(int(st[3]), int(total_rest), int(st[4]), int(st[5]), float(st[6]), float(st[7]), float(st[8]),
                          float(st[9]), float(st[10]), float(st[11]), int(bool(st[12])))


def forward(lib, params, st, training=True, to_chw=True, clamp=False, defer_color=False):
    """params: six tensors (means, scales, rotations, opacities, sh0, sh_rest); st: RasterizerSettings tuple.
    Returns (image, state, scratch, inputs) -- keep `scratch` and `inputs` alive for the backward."""
    p = [None if t is None else _copy_aligned(_f32(t)) for t in params]  # rotations None: covariance-input mode
    n = p[0].shape[0]
    total_rest = p[5].shape[1] if p[5].ndim > 1 else 0
    w2c, pos, bg = (_copy_aligned(_f32(t)) for t in st[:3])
    w, h = int(st[4]), int(st[5])
    image = _copy_aligned(np.full((3, h, w) if to_chw else (h, w, 3), np.nan, dtype=np.float32))
    s = settings_struct(st, total_rest)
    sc = Scratch()
    state = _capi.ForwardState()
    ptr = lambda a: None if a is None else a.ctypes.data
    if training and defer_color:
        rc = lib.fgs_forward_ex(sc.cb, 1, 2, 3, *[ptr(a) for a in p], ptr(w2c), ptr(pos), ptr(bg), ptr(image), n, C.byref(s), None,
                                C.byref(state), 1, None)
    elif training:
        rc = lib.fgs_forward(sc.cb, 1, 2, 3, *[ptr(a) for a in p], ptr(w2c), ptr(pos), ptr(bg), ptr(image), n, C.byref(s), None,
                             C.byref(state))
    else:
        rc = lib.fgs_inference(sc.cb, 1, 2, 3, *[ptr(a) for a in p], ptr(w2c), ptr(pos), ptr(bg), ptr(image), n, C.byref(s),
                               int(to_chw), int(clamp), None)
    assert rc == 0, lib.fgs_last_error()
    return image, state, sc, dict(p=p, w2c=w2c, pos=pos, bg=bg, s=s, n=n)


def backward(lib, image, grad_image, state, sc, inp, densification_info=None):
    p, n = inp['p'], inp['n']
    gi = _copy_aligned(_f32(grad_image))
    grads = [None if a is None else _copy_aligned(np.full(a.shape, np.nan, dtype=np.float32)) for a in p]
    scratch2d = _aligned(lib.fgs_grad2d_scratch_bytes(n))
    scratch2d[:] = 0xCD
    ptr = lambda a: None if a is None else a.ctypes.data
    dens = ptr(densification_info) if densification_info is not None else None
    rc = lib.fgs_backward(ptr(gi), ptr(image), ptr(p[0]), ptr(p[1]), ptr(p[2]), ptr(p[3]), ptr(p[5]), ptr(inp['w2c']),
                          ptr(inp['pos']), ptr(inp['bg']), ptr(sc.buf[1]), ptr(sc.buf[2]), ptr(sc.buf[3]),
                          *[ptr(g) for g in grads], ptr(scratch2d), dens, n, C.byref(inp['s']), C.byref(state), None)
    assert rc == 0, lib.fgs_last_error()
    return [None if g is None else torch.from_numpy(np.array(g)) for g in grads]


def backward_adam(lib, image, grad_image, state, sc, inp, moments_m, moments_v, steps, lrs, betas=(0.9, 0.999), eps=1e-15,
                  densification_info=None, defer_rest=False):
    """fgs_backward_adam (or fgs_backward_adam_ex with FGS_DEFER_SH_REST): updates inp['p'] (parameters) and the moment
    arrays in place.  On the emulator the side stream is synchronous: the deferred kernel has run when the call returns."""
    p, n = inp['p'], inp['n']
    gi = _copy_aligned(_f32(grad_image))
    scratch2d = _aligned(lib.fgs_grad2d_scratch_bytes(n))
    ptr = lambda a: a.ctypes.data
    arr = lambda ts: (C.c_void_p * 6)(*[ptr(t) for t in ts])
    dens = ptr(densification_info) if densification_info is not None else None
    args = (ptr(gi), ptr(image), arr(p), arr(moments_m), arr(moments_v), (C.c_int32 * 6)(*steps),
            (C.c_double * 6)(*lrs), betas[0], betas[1], eps, ptr(inp['w2c']), ptr(inp['pos']), ptr(inp['bg']),
            ptr(sc.buf[1]), ptr(sc.buf[2]), ptr(sc.buf[3]), ptr(scratch2d), dens, n, C.byref(inp['s']),
            C.byref(state), None)
    rc = lib.fgs_backward_adam_ex(*args, 1) if defer_rest else lib.fgs_backward_adam(*args)
    assert rc == 0, lib.fgs_last_error()


def adam_step(lib, grad, param, m, v, step, lr, betas=(0.9, 0.999), eps=1e-15):
    ptr = lambda a: a.ctypes.data
    rc = lib.fgs_adam_step(ptr(grad), ptr(param), ptr(m), ptr(v), param.size, int(step), float(lr), betas[0], betas[1], eps, None)
    assert rc == 0, lib.fgs_last_error()


def pruning_scores(lib, params, st, scores):
    p = [_copy_aligned(_f32(t)) for t in params]
    n = p[0].shape[0]
    w2c, pos, bg = (_copy_aligned(_f32(t)) for t in st[:3])
    w, h = int(st[4]), int(st[5])
    image = _copy_aligned(np.zeros((3, h, w), dtype=np.float32))
    s = settings_struct(st, p[5].shape[1])
    sc = Scratch()
    ptr = lambda a: a.ctypes.data
    rc = lib.fgs_pruning_scores(sc.cb, 1, 2, 3, *[ptr(a) for a in p], ptr(w2c), ptr(pos), ptr(bg), ptr(image), ptr(scores), n,
                                C.byref(s), None)
    assert rc == 0, lib.fgs_last_error()


def sort_pairs(lib, keys: np.ndarray, vals: np.ndarray, n: int, end_bit: int):
    """keys uint32 or uint16 of length n_max >= n; returns (sorted keys[:n], sorted vals[:n])."""
    n_max = keys.shape[0]
    k, v = _copy_aligned_any(keys), _copy_aligned_any(vals)
    k_alt, v_alt = _copy_aligned_any(np.zeros_like(keys)), _copy_aligned_any(np.zeros_like(vals))
    n_dev = _copy_aligned_any(np.array([n], dtype=np.uint32))
    sb = lib.fgs_sort_scratch_bytes(n_max, end_bit)
    scratch = _aligned(sb)
    scratch[:] = 0xCD
    in_alt = C.c_int32(0)
    fn = lib.fgs_sort_pairs_u32 if keys.dtype == np.uint32 else lib.fgs_sort_pairs_u16
    rc = fn(k.ctypes.data, k_alt.ctypes.data, v.ctypes.data, v_alt.ctypes.data, n_dev.ctypes.data, n_max, end_bit,
            scratch.ctypes.data, sb, C.byref(in_alt), None)
    assert rc == 0, lib.fgs_last_error()
    ks, vs = (k_alt, v_alt) if in_alt.value else (k, v)
    return np.array(ks[:n]), np.array(vs[:n])


def _copy_aligned_any(a: np.ndarray) -> np.ndarray:
    out = _aligned(max(a.nbytes, 16)).view(a.dtype)[:a.size].reshape(a.shape)
    out[...] = a
    return out
