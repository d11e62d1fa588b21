"""Operators of the reference API around the rasterizer: same names and argument lists as
/root/reference/FasterGSCudaBackend/FasterGSCudaBackend/torch_bindings/filter3d.py:6-35 and
densification.py:6-22, backed by fgs_b200._C."""
import torch

from .. import _C


def update_3d_filter(positions, w2c, filter_3d, visibility_mask, width, height, focal_x, focal_y, center_x, center_y,
                     near_plane, clipping_tolerance, distance2filter) -> None:
    _C.update_3d_filter(positions, w2c, filter_3d, visibility_mask, width, height, focal_x, focal_y, center_x, center_y,
                        near_plane, clipping_tolerance, distance2filter)


def relocation_adjustment(old_opacities, old_scales, n_samples_per_primitive):
    return _C.relocation_adjustment(old_opacities, old_scales, n_samples_per_primitive)


def add_noise(raw_scales, raw_rotations, raw_opacities, means, current_lr) -> None:
    random_samples = torch.randn_like(means)  # drawn by torch, as in the reference (densification.py:21)
    _C.add_noise(raw_scales, raw_rotations, raw_opacities, random_samples, means, current_lr)
