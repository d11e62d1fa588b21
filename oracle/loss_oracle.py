"""CPU ORACLE (test infrastructure only -- never imported by the product path) of the training loss.

Restates /root/reference/Loss.py:15-16,26-33: ``LAMBDA_L1 * l1_loss(image, target) + LAMBDA_DSSIM * fused_dssim(image,
target)``.  ``fused_dssim`` is NeRFICG's ``Optim.Losses.DSSIM.fused_dssim`` (Loss.py:8) -- NOT in the reference tree and
not version-pinned, so this file restates the published algorithm it wraps (fused-ssim, as used by
3DGS-derived trainers): SSIM per channel with an 11x11 Gaussian window (sigma 1.5), zero padding ("same" convolution),
C1 = 0.01**2, C2 = 0.03**2, averaged over all 3*H*W values; DSSIM = 1 - mean SSIM.  **Parity unpinned** for the DSSIM
term: there is no reference output to check it against; what IS checked is (a) this restatement against an independent
dense (non-separable, explicitly padded) evaluation and (b) the analytic gradient of the CUDA kernels against autograd
through this function.  The L1 term is torch's own ``l1_loss`` and needs no pin.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

C1 = 0.01 ** 2
C2 = 0.03 ** 2


def gaussian_window(dtype=torch.float64) -> torch.Tensor:
    """11 taps, sigma 1.5, normalised to sum 1 (float32 values as the kernel uses them, then widened)."""
    k = torch.arange(11, dtype=torch.float64) - 5.0
    g = torch.exp(-k * k / (2.0 * 1.5 * 1.5))
    return (g / g.sum()).to(torch.float32).to(dtype)


def _blur(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """Separable zero-padded 11x11 Gaussian blur of a (C,H,W) tensor, per channel."""
    c = x.shape[0]
    kh = g.reshape(1, 1, 1, 11).expand(c, 1, 1, 11)
    kv = g.reshape(1, 1, 11, 1).expand(c, 1, 11, 1)
    y = F.conv2d(x[None], kh, padding=(0, 5), groups=c)
    return F.conv2d(y, kv, padding=(5, 0), groups=c)[0]


def ssim_map(image: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    g = gaussian_window(image.dtype)
    mu1, mu2 = _blur(image, g), _blur(target, g)
    s11 = _blur(image * image, g) - mu1 * mu1
    s22 = _blur(target * target, g) - mu2 * mu2
    s12 = _blur(image * target, g) - mu1 * mu2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))


def ssim_map_dense(image: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Independent evaluation for the self-check: explicit zero padding, full 11x11 window, unfold."""
    g = gaussian_window(image.dtype)
    w2 = torch.outer(g, g).reshape(-1)

    def blur(x):
        c, h, w = x.shape
        patches = F.unfold(F.pad(x[:, None], (5, 5, 5, 5)), kernel_size=11)  # (C, 121, H*W)
        return (patches * w2[None, :, None]).sum(1).reshape(c, h, w)

    mu1, mu2 = blur(image), blur(target)
    s11, s22, s12 = blur(image * image) - mu1 ** 2, blur(target * target) - mu2 ** 2, blur(image * target) - mu1 * mu2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s11 + s22 + C2))


def loss_terms(image: torch.Tensor, target: torch.Tensor):
    """(L1, DSSIM) of Loss.py:15-16; target may be uint8 (value / 255)."""
    if target.dtype == torch.uint8:
        target = target.to(torch.float32).div(255.0)
    target = target.to(image.dtype)
    l1 = (image - target).abs().mean()
    dssim = 1.0 - ssim_map(image, target).mean()
    return l1, dssim


def loss_and_grad(image: torch.Tensor, target: torch.Tensor, lambda_l1: float, lambda_dssim: float,
                  grad_loss: float = 1.0, dtype=torch.float64):
    """Returns (loss, l1, dssim, dloss/dimage) evaluated in `dtype` with autograd."""
    x = image.detach().to(dtype).clone().requires_grad_(True)
    t = target if target.dtype == torch.uint8 else target.detach().to(dtype)
    l1, ds = loss_terms(x, t)
    loss = lambda_l1 * l1 + (lambda_dssim * ds if lambda_dssim != 0.0 else 0.0)
    (g,) = torch.autograd.grad(loss * grad_loss, x)
    return loss.detach(), l1.detach(), ds.detach(), g
