"""CPU oracle for the three small operators around the rasterizer.  TEST INFRASTRUCTURE ONLY (see
oracle/torch_oracle.py for the rules).  Restated from the reference:

  update_3d_filter ....... FasterGSCudaBackend/FasterGSCudaBackend/filter3d/src/filter3d.cu:9-86
  relocation_adjustment .. .../densification/include/kernels_mcmc.cuh:10-57 (3DGS-MCMC Eq. 9, <= 50 samples)
  add_noise .............. .../densification/include/kernels_mcmc.cuh:69-127

Pinned against tests/golden/aux_ops.npz (dumped from the compiled reference on a B200)."""
from __future__ import annotations

import math

import torch

MAX_SAMPLES = 50  # densification_config.h:9


def update_3d_filter(positions, w2c, filter_3d, visibility_mask, width, height, focal_x, focal_y, center_x, center_y,
                     near_plane, clipping_tolerance, distance2filter) -> None:
    """In place on filter_3d (float) and visibility_mask (bool)."""
    bf = clipping_tolerance + 0.5
    ox, oy = center_x - 0.5 * width, center_y - 0.5 * height
    left, right = (-bf * width - ox) / focal_x, (bf * width - ox) / focal_x
    top, bottom = (-bf * height - oy) / focal_y, (bf * height - oy) / focal_y
    W = w2c[:3, :3].to(positions.dtype)
    t = w2c[:3, 3].to(positions.dtype)
    cam = positions @ W.T + t
    x, y, z = cam[:, 0], cam[:, 1], cam[:, 2]
    new = distance2filter * z
    ok = (z >= near_plane) & (x >= left * z) & (x <= right * z) & (y >= top * z) & (y <= bottom * z) & ~(filter_3d < new)
    filter_3d[ok] = new[ok].to(filter_3d.dtype)
    visibility_mask[ok] = True


def relocation_coefficients() -> torch.Tensor:
    """c[n][k] = (-1)^k C(n,k) / sqrt(k+1), built in double like kernels_mcmc.cuh:13-26."""
    c = torch.zeros(MAX_SAMPLES, MAX_SAMPLES, dtype=torch.float64)
    for n in range(MAX_SAMPLES):
        binom, sign = 1.0, 1.0
        for k in range(n + 1):
            c[n, k] = binom * sign / math.sqrt(k + 1)
            binom *= (n - k) / (k + 1)
            sign = -sign
    return c.float()


def relocation_adjustment(old_opacities, old_scales, n_samples_per_primitive):
    """Returns (new_opacities (n,1), new_scales (n,3)); float32 arithmetic in the reference's summation
    order (the alternating binomial series cancels heavily, so the order is part of the definition)."""
    o = old_opacities.reshape(-1).float()
    ns = n_samples_per_primitive.reshape(-1).clamp(1, MAX_SAMPLES).long()
    o_new = (1.0 - torch.pow(1.0 - o, 1.0 / ns.float())).float()
    c = relocation_coefficients()
    denom = torch.zeros_like(o)
    for n in range(MAX_SAMPLES):
        live = ns > n
        if not bool(live.any()):
            break
        pw = o_new.clone()
        for k in range(n + 1):
            denom = torch.where(live, denom + c[n, k] * pw, denom)
            pw = pw * o_new
    f = o / denom
    return o_new.reshape(-1, 1), f[:, None] * old_scales.float()


def add_noise(raw_scales, raw_rotations, raw_opacities, random_samples, means, current_lr) -> None:
    """In place on means."""
    from oracle.torch_oracle import quaternion_to_rotation
    R, norm_sq = quaternion_to_rotation(raw_rotations)
    var = torch.exp(2.0 * raw_scales)
    cov = R @ torch.diag_embed(var) @ R.transpose(1, 2)
    noise = torch.einsum('nij,nj->ni', cov, random_samples)
    o = torch.sigmoid(raw_opacities.reshape(-1))
    gate = 1.0 / (1.0 + torch.exp(100.0 * o - 0.5))
    upd = (current_lr * gate)[:, None] * noise
    ok = norm_sq >= 1e-8
    means[ok] += upd[ok].to(means.dtype)
