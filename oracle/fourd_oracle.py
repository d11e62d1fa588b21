"""CPU ORACLE (test infrastructure only -- never imported by the product path) of the time slice of 4D Gaussians
(BASELINE.json configs[4]).

The reference's `FasterGS4D` branch (README.md:39) is NOT in the tree; it extends the rasterizer to the 4D Gaussians of
Yang et al., "Real-time Photorealistic Dynamic Scene Representation and Rendering with 4D Gaussian Splatting" (ICLR 2024).
**Parity unpinned**: this file restates the published construction, nothing of the reference's implementation can be
checked against it.

A 4D Gaussian has a mean (x, y, z, t), four log-scales and two unnormalised quaternions q_l, q_r; its 4x4 rotation is the
product of a left- and a right-isoclinic rotation, R = L(q_l / |q_l|) R(q_r / |q_r|), and Sigma4 = R diag(exp(s))^2 R^T.
Conditioning on the time t gives an ordinary 3D Gaussian,

    mean3 = mu_xyz + Sigma4[:3, 3] (t - mu_t) / Sigma4[3, 3]
    cov3  = Sigma4[:3, :3] - Sigma4[:3, 3] Sigma4[3, :3] / Sigma4[3, 3]

whose opacity is scaled by the temporal marginal exp(-(t - mu_t)^2 / (2 Sigma4[3, 3])); the result is rendered by the
rasterizer in covariance-input mode.  The colour is the static SH of the 3D model (the paper's 4D spherindrical harmonics
are not part of this restatement)."""
from __future__ import annotations

import torch


def left_isoclinic(q):
    a, b, c, d = q.unbind(1)
    return torch.stack([a, -b, -c, -d,
                        b, a, -d, c,
                        c, d, a, -b,
                        d, -c, b, a], dim=1).reshape(-1, 4, 4)


def right_isoclinic(q):
    p, q_, r, s = q.unbind(1)
    return torch.stack([p, q_, r, s,
                        -q_, p, -s, r,
                        -r, s, p, -q_,
                        -s, -r, q_, p], dim=1).reshape(-1, 4, 4)


def covariance_4d(scales4, rot_l, rot_r):
    ql = rot_l / rot_l.norm(dim=1, keepdim=True)
    qr = rot_r / rot_r.norm(dim=1, keepdim=True)
    R = left_isoclinic(ql) @ right_isoclinic(qr)
    M = R * torch.exp(scales4)[:, None, :]
    return M @ M.transpose(1, 2)


def slice_4d(means4, scales4, rot_l, rot_r, opacities, t):
    """-> (means3 (N,3), cov3 (N,6) = (xx, xy, xz, yy, yz, zz), opacity logits (N,1)) of the 3D Gaussians at time t."""
    sigma = covariance_4d(scales4, rot_l, rot_r)
    A, b, c = sigma[:, :3, :3], sigma[:, :3, 3], sigma[:, 3, 3]
    dt = t - means4[:, 3]
    means3 = means4[:, :3] + b * (dt / c)[:, None]
    cov = A - b[:, :, None] * b[:, None, :] / c[:, None, None]
    cov3 = torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], dim=1)
    marginal = torch.exp(-0.5 * dt * dt / c)
    o = torch.sigmoid(opacities[:, 0]) * marginal
    return means3, cov3, torch.logit(o)[:, None]


def render_4d(means4, scales4, rot_l, rot_r, opacities, sh0, sh_rest, t, settings):
    from oracle import torch_oracle as O
    means3, cov3, logits = slice_4d(means4, scales4, rot_l, rot_r, opacities, t)
    return O.render(means3, None, None, logits, sh0, sh_rest, settings, cov3d=cov3)
