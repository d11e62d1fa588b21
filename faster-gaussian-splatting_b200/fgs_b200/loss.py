"""Training loss fused with its gradient: the step between `diff_rasterize` and `loss.backward()`.

Mirrors /root/reference/Loss.py:15-16,26-33 as called by Trainer.py:191-196:

    loss = lambda_l1 * l1_loss(image, target) + lambda_dssim * fused_dssim(image, target)

`l1_dssim_loss(image, target, lambda_l1, lambda_dssim)` is the autograd form (a scalar, differentiable w.r.t. `image`);
`l1_dssim_loss_and_grad` is the one-call form for a trainer that backpropagates right away
(`image.backward(grad_image)`): it returns the three loss scalars (weighted loss, L1, DSSIM) and d loss / d image.
The target may be float32 in [0, 1] or uint8 (value / 255: the dataset's storage format, no conversion pass).
All computation happens in csrc/loss.cu behind the C ABI (fgs_loss_*); there is no torch fallback.
DSSIM = 1 - mean SSIM (11x11 Gaussian window, sigma 1.5, zero padding) as in NeRFICG's fused_dssim, which is not in the
reference tree: its parity is pinned by oracle/loss_oracle.py only."""
from __future__ import annotations

import torch

from . import _capi

_lib = _capi.load()


def _check(image: torch.Tensor, target: torch.Tensor):
    if not (_capi.on_device(image) and image.dtype == torch.float32 and image.dim() == 3 and image.shape[0] == 3):
        raise RuntimeError("'image' must be a CUDA float tensor of shape (3, H, W).")
    if target.shape != image.shape or target.device != image.device or target.dtype not in (torch.float32, torch.uint8):
        raise RuntimeError("'target' must be a float32 or uint8 tensor of the image's shape on the image's device.")
    image, target = image.contiguous(), target.contiguous()
    # the kernels use 16-byte vector accesses: a view at an odd offset (e.g. images[k] of a stack of odd-sized images)
    # is copied once
    if image.data_ptr() % 16:
        image = image.clone()
    if target.data_ptr() % 16:
        target = target.clone()
    return image, target


def _scratch(image: torch.Tensor) -> torch.Tensor:
    _, h, w = image.shape
    return torch.empty(_lib.fgs_loss_scratch_bytes(w, h), dtype=torch.uint8, device=image.device)


def _stream() -> int:
    return _capi.current_stream()


@torch.no_grad()
def l1_dssim_loss_and_grad(image, target, lambda_l1: float, lambda_dssim: float):
    """-> (terms (3,) = [loss, l1, dssim] on the device, grad_image (3,H,W)) with dL/dloss = 1; no host sync."""
    image, target = _check(image, target)
    _, h, w = image.shape
    with _capi.device_guard(image.device):
        scratch = _scratch(image)
        terms = torch.empty(3, dtype=torch.float32, device=image.device)
        grad = torch.empty_like(image)
        rc = _lib.fgs_loss_forward_backward(image.data_ptr(), target.data_ptr(), int(target.dtype == torch.uint8), w, h,
                                            float(lambda_l1), float(lambda_dssim), scratch.data_ptr(), terms.data_ptr(),
                                            grad.data_ptr(), _stream())
        _capi.check(rc, 'loss_forward_backward')
    return terms, grad


class _L1DSSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, target, lambda_l1, lambda_dssim):
        image_c, target_c = _check(image, target)
        _, h, w = image_c.shape
        with _capi.device_guard(image_c.device):
            scratch = _scratch(image_c)
            terms = torch.empty(3, dtype=torch.float32, device=image_c.device)
            rc = _lib.fgs_loss_forward(image_c.data_ptr(), target_c.data_ptr(), int(target_c.dtype == torch.uint8), w, h,
                                       float(lambda_l1), float(lambda_dssim), scratch.data_ptr(), terms.data_ptr(),
                                       _stream())
            _capi.check(rc, 'loss_forward')
        ctx.save_for_backward(image_c, target_c, scratch)
        ctx.lambdas = (float(lambda_l1), float(lambda_dssim))
        ctx.mark_non_differentiable(terms)
        return terms[0].clone(), terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_terms):
        image, target, scratch = ctx.saved_tensors
        _, h, w = image.shape
        with _capi.device_guard(image.device):
            gl = grad_loss.to(device=image.device, dtype=torch.float32).reshape(1).contiguous()
            grad = torch.empty_like(image)
            rc = _lib.fgs_loss_backward(image.data_ptr(), target.data_ptr(), int(target.dtype == torch.uint8), w, h,
                                        ctx.lambdas[0], ctx.lambdas[1], scratch.data_ptr(), gl.data_ptr(), grad.data_ptr(),
                                        _stream())
            _capi.check(rc, 'loss_backward')
        return grad, None, None, None


def l1_dssim_loss(image, target, lambda_l1: float = 0.8, lambda_dssim: float = 0.2, return_terms: bool = False):
    """Scalar loss, differentiable w.r.t. `image` (defaults: fastergs_garden.yaml LOSS section).  With
    `return_terms` also the (3,) tensor [loss, l1, dssim] for logging."""
    loss, terms = _L1DSSIM.apply(image, target, lambda_l1, lambda_dssim)
    return (loss, terms) if return_terms else loss
