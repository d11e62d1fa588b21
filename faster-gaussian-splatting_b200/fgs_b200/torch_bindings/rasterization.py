"""Python operator API of the rasterizer: same names, arguments and behaviour as the reference's
torch_bindings/rasterization.py (RasterizerSettings :8-38, _Rasterize :41-110, diff_rasterize
:113-132, rasterize :135-156, update_pruning_scores :159-178), backed by fgs_b200._C.

Additional entry point `diff_rasterize_fused` gives the FasterGSFused behaviour (README.md:37):
the backward pass applies the FusedAdam update in the same kernel and returns no gradients."""
from typing import Any, NamedTuple

import torch
from torch.autograd.function import once_differentiable

from .. import _C


class RasterizerSettings(NamedTuple):
    """Field names, order and types are the drop-in boundary (the reference builds this tuple in Renderer.py:19-43 and
    unpacks it positionally into `_C.forward`); what each field means to the kernels here:"""
    w2c: torch.Tensor            # (>=3, 4) row-major world -> camera matrix on the device; rows 0..2 are read, +z looks forward
    cam_position: torch.Tensor   # (3,) camera centre in world coordinates, device tensor (SH view directions)
    bg_color: torch.Tensor       # (3,) RGB added behind the last Gaussian (image += T_final * bg), device tensor
    active_sh_bases: int         # 1, 4, 9 or 16: SH bases evaluated (degree 0..3); higher coefficients get zero gradient
    width: int                   # image width in pixels (tiles of 16 x 12 pixels, ragged edges allowed)
    height: int                  # image height in pixels
    focal_x: float               # pixels: x_image = focal_x * x / z + center_x
    focal_y: float               # pixels: y_image = focal_y * y / z + center_y
    center_x: float              # principal point, absolute pixel coordinate, +x to the right
    center_y: float              # principal point, absolute pixel coordinate, +y downwards
    near_plane: float            # Gaussians with camera depth z < near_plane are culled
    far_plane: float             # ... and those with z > far_plane
    proper_antialiasing: bool    # 0.1 px dilation with opacity compensation instead of the plain 0.3 px dilation

    def as_tuple(self) -> tuple:
        return tuple(self)


_forward_kwargs = None


def set_forward_kwargs(provider) -> None:
    """`provider()` -> keyword arguments of `_C.forward_ex` for every following `diff_rasterize` forward (None: plain
    forward).  Used by the view-parallel optimizer whose sh_rest parameters arrive from the peer GPUs while the forward's
    geometry pass already runs (fgs_b200.parallel.PeerAdam(compact=True)): {'defer_color': True, 'color_event': ...}."""
    global _forward_kwargs
    _forward_kwargs = provider


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                densification_info, rasterizer_settings: RasterizerSettings) -> torch.Tensor:
        kwargs = _forward_kwargs() if _forward_kwargs is not None else {}
        (image, primitive_buffers, tile_buffers, instance_buffers, bucket_buffers,
         n_instances, n_buckets, selector) = _C.forward_ex(
            means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
            *rasterizer_settings.as_tuple(), **kwargs)
        ctx.rasterizer_settings = rasterizer_settings
        ctx.buffer_state = (n_instances, n_buckets, selector)
        ctx.save_for_backward(image, means, scales, rotations, opacities, sh_coefficients_rest,
                              primitive_buffers, tile_buffers, instance_buffers, bucket_buffers)
        ctx.densification_info = densification_info
        ctx.mark_non_differentiable(densification_info)
        return image

    @staticmethod
    @once_differentiable
    def backward(ctx: Any, grad_image: torch.Tensor):
        grads = _C.backward(ctx.densification_info, grad_image, *ctx.saved_tensors,
                            *ctx.rasterizer_settings.as_tuple(), *ctx.buffer_state)
        return (*grads, None, None)


class _RasterizeFused(torch.autograd.Function):
    """Forward identical to _Rasterize; backward = gradient + Adam update in one kernel.  The six
    parameters must be leaves managed by `optimizer` (one tensor per group).  No gradients are
    returned, so a following `optimizer.step()` is a no-op for them (adam.py:17-18)."""

    @staticmethod
    def forward(ctx: Any, means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                densification_info, rasterizer_settings: RasterizerSettings, optimizer, defer_sh: bool = False) -> torch.Tensor:
        (image, primitive_buffers, tile_buffers, instance_buffers, _bucket_buffers,
         n_instances, _n_buckets, _selector) = _C.forward_ex(
            means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
            *rasterizer_settings.as_tuple(), defer_color=bool(defer_sh))
        ctx.defer_sh = bool(defer_sh)
        ctx.rasterizer_settings = rasterizer_settings
        ctx.n_instances = n_instances
        ctx.params = (means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest)
        ctx.optimizer = optimizer
        ctx.save_for_backward(image, primitive_buffers, tile_buffers, instance_buffers)
        ctx.densification_info = densification_info
        ctx.mark_non_differentiable(densification_info)
        return image

    @staticmethod
    @once_differentiable
    def backward(ctx: Any, grad_image: torch.Tensor):
        image, primitive_buffers, tile_buffers, instance_buffers = ctx.saved_tensors
        opt = ctx.optimizer
        params = [p.detach() for p in ctx.params]
        groups = [opt.group_of(p) for p in ctx.params]
        states = [opt._state_for(p) for p in ctx.params]
        betas, eps = groups[0]['betas'], groups[0]['eps']
        for g in groups:
            if g['betas'] != betas or g['eps'] != eps:
                raise RuntimeError('diff_rasterize_fused: all six param groups must share betas and eps')
        for s in states:
            s['step'] += 1
        _C.backward_adam(ctx.densification_info, grad_image, image, params,
                         [s['exp_avg'] for s in states], [s['exp_avg_sq'] for s in states],
                         [s['step'] for s in states], [g['lr'] for g in groups], betas[0], betas[1], eps,
                         primitive_buffers, tile_buffers, instance_buffers,
                         *ctx.rasterizer_settings.as_tuple(), ctx.n_instances, defer_rest=ctx.defer_sh)
        return (None,) * 10


def diff_rasterize(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                   densification_info, rasterizer_settings: RasterizerSettings) -> torch.Tensor:
    return _Rasterize.apply(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                            densification_info, rasterizer_settings)


def diff_rasterize_fused(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                         densification_info, rasterizer_settings: RasterizerSettings, optimizer,
                         defer_sh: bool = False) -> torch.Tensor:
    """`defer_sh=True`: the deferred SH schedule (include/fgs_b200.h) -- identical parameters and moments after the
    step, but the Adam update of sh_coefficients_rest runs on a side stream under the NEXT forward's geometry pass and
    sorts.  Between iterations sh_coefficients_rest and its moments may still be in flight: call
    `fgs_b200.join_deferred()` before reading or writing them with anything but this package's operators (checkpoints,
    densification with plain torch ops, ...)."""
    return _RasterizeFused.apply(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                                 densification_info, rasterizer_settings, optimizer, defer_sh)


def rasterize(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
              rasterizer_settings: RasterizerSettings, to_chw: bool, clamp_output: bool = True) -> torch.Tensor:
    return _C.inference(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                        *rasterizer_settings.as_tuple(), to_chw, clamp_output)


def update_pruning_scores(scores, means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                          rasterizer_settings: RasterizerSettings) -> torch.Tensor:
    return _C.pruning_scores(scores, means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest,
                             *rasterizer_settings.as_tuple())
