"""Synthetic scenes and cameras of the benchmark shapes (SURVEY.md §8d `make_scene`).

Pure torch, CPU-generated and seed-deterministic so the GPU box, this container and the
oracle all see identical inputs. Parameter layouts are the reference's `nn.Parameter`s
(/root/reference/Model.py:226-231): means (N,3), raw log-scales (N,3), raw quaternions (N,4)
in (w,x,y,z) order, opacity logits (N,1), sh0 (N,1,3), sh_rest (N,15,3).
"""
from __future__ import annotations

import math
from typing import NamedTuple

import torch

SH_C0 = 0.28209479177387814
SIGMA_AT_1M = 0.0045  # calibrated with oracle/calibrate_scene.py


class Scene(NamedTuple):
    means: torch.Tensor
    scales: torch.Tensor
    rotations: torch.Tensor
    opacities: torch.Tensor
    sh_coefficients_0: torch.Tensor
    sh_coefficients_rest: torch.Tensor

    def to(self, device) -> 'Scene':
        return Scene(*[t.to(device).contiguous() for t in self])

    def params(self) -> list[torch.Tensor]:
        return list(self)


class Camera(NamedTuple):
    w2c: torch.Tensor           # (4,4) row-major, +z forward
    position: torch.Tensor      # (3,)
    width: int
    height: int
    focal_x: float
    focal_y: float
    center_x: float
    center_y: float
    near_plane: float
    far_plane: float


def default_log_scale_mean(n: int, width: int) -> float:
    """Scale calibration: keeps instances-per-visible-Gaussian in the trained-scene range
    (SURVEY.md §8d asks for I/V in [3,6]) across N and resolution.

    The projected footprint radius in pixels is ~ 3.33 * sigma * focal / depth; with
    focal = 0.75*width and depth ~ 4, holding the total covered area (N * radius^2) per
    pixel roughly constant gives sigma ~ sqrt(1/N) at fixed resolution."""
    sigma = SIGMA_AT_1M * math.sqrt(1.0e6 / max(n, 1))
    return math.log(sigma)


def make_scene(n: int, seed: int = 0, log_scale_mean: float | None = None, width: int = 1920) -> Scene:
    g = torch.Generator(device='cpu').manual_seed(seed)
    if log_scale_mean is None:
        log_scale_mean = default_log_scale_mean(n, width)
    means = (torch.randn(n, 3, generator=g) * 1.5).clamp_(-6.0, 6.0)
    scales = torch.randn(n, 3, generator=g) * 0.7 + log_scale_mean
    rotations = torch.randn(n, 4, generator=g)
    opacities = (torch.randn(n, 1, generator=g) * 2.0).clamp_(-6.0, 6.0)
    sh0 = (torch.rand(n, 1, 3, generator=g) - 0.5) / SH_C0
    sh_rest = torch.randn(n, 15, 3, generator=g) * 0.05
    return Scene(means.contiguous(), scales.contiguous(), rotations.contiguous(),
                 opacities.contiguous(), sh0.contiguous(), sh_rest.contiguous())


def make_camera(k: int, width: int, height: int, n_views: int = 8,
                radius: float = 4.0, cam_height: float = 1.0) -> Camera:
    """Camera k of a ring looking at the origin."""
    az = 2.0 * math.pi * (k % n_views) / n_views
    pos = torch.tensor([radius * math.cos(az), cam_height, radius * math.sin(az)], dtype=torch.float64)
    fwd = -pos / pos.norm()
    up_hint = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)  # +y of the image points down
    right = torch.linalg.cross(up_hint, fwd)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    rot = torch.stack([right, down, fwd])  # rows = camera axes in world coordinates
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = rot
    w2c[:3, 3] = -rot @ pos
    return Camera(w2c.float().contiguous(), pos.float().contiguous(), width, height,
                  0.75 * width, 0.75 * width, width / 2.0, height / 2.0, 0.2, 1.0e4)


def make_target(width: int, height: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device='cpu').manual_seed(seed + 100)
    return torch.rand(3, height, width, generator=g)


def make_grad_image(width: int, height: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device='cpu').manual_seed(seed + 200)
    return torch.randn(3, height, width, generator=g)
