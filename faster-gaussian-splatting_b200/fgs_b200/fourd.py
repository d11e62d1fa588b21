"""4D Gaussians (BASELINE.json configs[4]): time slice + the rasterizer's covariance-input mode.

The reference's FasterGS4D branch (README.md:39) is not in the tree; this follows Yang et al. (ICLR 2024), restated in
oracle/fourd_oracle.py -- parity unpinned.  Operators (all backed by the C ABI, csrc/slice4d.cu and the covariance-input
mode of fgs_forward / fgs_backward; no torch fallback):

    means3, cov3, logits = slice_4d(means4, scales4, rot_l, rot_r, opacities, t)         # differentiable
    image = diff_rasterize_cov3d(means3, cov3, logits, sh0, sh_rest, densification_info, settings)
    image = diff_rasterize_4d(means4, scales4, rot_l, rot_r, opacities, sh0, sh_rest, t, densification_info, settings)
    image = rasterize_4d(..., t, settings, to_chw, clamp_output=True)                     # inference

means4 (N,4) = (x, y, z, t); scales4 (N,4) log-scales; rot_l / rot_r (N,4) unnormalised quaternions of the left / right
isoclinic rotation; opacities (N,1) logits; cov3 (N,6) = (xx, xy, xz, yy, yz, zz).  The colour is the static SH of the 3D
model.  `settings` is the same RasterizerSettings as for diff_rasterize."""
from __future__ import annotations

from typing import Any

import torch
from torch.autograd.function import once_differentiable

from . import _C, _capi
from .torch_bindings.rasterization import RasterizerSettings

_lib = _capi.load()


def _check(t: torch.Tensor, name: str, cols: int) -> torch.Tensor:
    if not (_capi.on_device(t) and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 2 and t.shape[1] == cols):
        raise RuntimeError(f"'{name}' must be a contiguous CUDA float tensor of shape (N, {cols}).")
    return t


class _Slice4D(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, means4, scales4, rot_l, rot_r, opacities, t: float):
        for x, name, cols in ((means4, 'means4', 4), (scales4, 'scales4', 4), (rot_l, 'rot_l', 4), (rot_r, 'rot_r', 4),
                              (opacities, 'opacities', 1)):
            _check(x, name, cols)
        n, dev = means4.shape[0], means4.device
        means3 = torch.empty((n, 3), dtype=torch.float32, device=dev)
        cov3 = torch.empty((n, 6), dtype=torch.float32, device=dev)
        logits = torch.empty((n, 1), dtype=torch.float32, device=dev)
        with _capi.device_guard(dev):
            rc = _lib.fgs_slice4d_forward(means4.data_ptr(), scales4.data_ptr(), rot_l.data_ptr(), rot_r.data_ptr(),
                                          opacities.data_ptr(), float(t), n, means3.data_ptr(), cov3.data_ptr(), logits.data_ptr(),
                                          _capi.current_stream())
            _capi.check(rc, 'slice4d_forward')
        ctx.save_for_backward(means4, scales4, rot_l, rot_r, opacities)
        ctx.t = float(t)
        return means3, cov3, logits

    @staticmethod
    @once_differentiable
    def backward(ctx: Any, g_means3, g_cov3, g_logits):
        means4, scales4, rot_l, rot_r, opacities = ctx.saved_tensors
        n, dev = means4.shape[0], means4.device
        zeros = lambda g, shape: torch.zeros(shape, dtype=torch.float32, device=dev) if g is None else g.to(torch.float32).contiguous()
        g_means3, g_cov3, g_logits = zeros(g_means3, (n, 3)), zeros(g_cov3, (n, 6)), zeros(g_logits, (n, 1))
        grads = [torch.empty_like(x) for x in (means4, scales4, rot_l, rot_r, opacities)]
        with _capi.device_guard(dev):
            rc = _lib.fgs_slice4d_backward(means4.data_ptr(), scales4.data_ptr(), rot_l.data_ptr(), rot_r.data_ptr(),
                                           opacities.data_ptr(), ctx.t, n, g_means3.data_ptr(), g_cov3.data_ptr(), g_logits.data_ptr(),
                                           *[g.data_ptr() for g in grads], _capi.current_stream())
            _capi.check(rc, 'slice4d_backward')
        return (*grads, None)


def slice_4d(means4, scales4, rot_l, rot_r, opacities, t: float):
    """The 3D Gaussians at time t: (means3 (N,3), cov3 (N,6), opacity logits (N,1))."""
    return _Slice4D.apply(means4, scales4, rot_l, rot_r, opacities, t)


class _RasterizeCov3D(torch.autograd.Function):
    """diff_rasterize with the 3D covariance given directly (rotations == None on the native side)."""

    @staticmethod
    def forward(ctx: Any, means, cov3, opacities, sh_coefficients_0, sh_coefficients_rest, densification_info,
                rasterizer_settings: RasterizerSettings) -> torch.Tensor:
        _check(cov3, 'cov3', 6)
        (image, primitive_buffers, tile_buffers, instance_buffers, bucket_buffers, n_instances, n_buckets, selector) = _C.forward(
            means, cov3, None, opacities, sh_coefficients_0, sh_coefficients_rest, *rasterizer_settings.as_tuple())
        ctx.rasterizer_settings = rasterizer_settings
        ctx.buffer_state = (n_instances, n_buckets, selector)
        ctx.save_for_backward(image, means, cov3, opacities, sh_coefficients_rest, primitive_buffers, tile_buffers,
                              instance_buffers, bucket_buffers)
        ctx.densification_info = densification_info
        ctx.mark_non_differentiable(densification_info)
        return image

    @staticmethod
    @once_differentiable
    def backward(ctx: Any, grad_image: torch.Tensor):
        image, means, cov3, opacities, sh_rest, prim, tile, inst, bucket = ctx.saved_tensors
        g_means, g_cov3, _, g_opac, g_sh0, g_rest = _C.backward(
            ctx.densification_info, grad_image, image, means, cov3, None, opacities, sh_rest, prim, tile, inst, bucket,
            *ctx.rasterizer_settings.as_tuple(), *ctx.buffer_state)
        return g_means, g_cov3, g_opac, g_sh0, g_rest, None, None


def diff_rasterize_cov3d(means, cov3, opacities, sh_coefficients_0, sh_coefficients_rest, densification_info,
                         rasterizer_settings: RasterizerSettings) -> torch.Tensor:
    return _RasterizeCov3D.apply(means, cov3, opacities, sh_coefficients_0, sh_coefficients_rest, densification_info,
                                 rasterizer_settings)


def diff_rasterize_4d(means4, scales4, rot_l, rot_r, opacities, sh_coefficients_0, sh_coefficients_rest, t: float,
                      densification_info, rasterizer_settings: RasterizerSettings) -> torch.Tensor:
    """Training forward of a 4D model at time t; differentiable w.r.t. the seven parameter tensors."""
    means3, cov3, logits = slice_4d(means4, scales4, rot_l, rot_r, opacities, t)
    return diff_rasterize_cov3d(means3, cov3, logits, sh_coefficients_0, sh_coefficients_rest, densification_info,
                                rasterizer_settings)


@torch.no_grad()
def rasterize_4d(means4, scales4, rot_l, rot_r, opacities, sh_coefficients_0, sh_coefficients_rest, t: float,
                 rasterizer_settings: RasterizerSettings, to_chw: bool, clamp_output: bool = True) -> torch.Tensor:
    """Inference render of a 4D model at time t ((3,H,W) if to_chw else (H,W,3))."""
    means3, cov3, logits = _Slice4D.apply(means4, scales4, rot_l, rot_r, opacities, t)
    return _C.inference(means3, cov3, None, logits, sh_coefficients_0, sh_coefficients_rest, *rasterizer_settings.as_tuple(),
                        to_chw, clamp_output)
