"""fgs_b200: Blackwell-native (sm_100a) drop-in for the FasterGSCudaBackend operator API.

    from fgs_b200 import diff_rasterize, rasterize, RasterizerSettings, FusedAdam

The eight public names of the reference backend (/root/reference/FasterGSCudaBackend/__init__.py:14-18)
plus `diff_rasterize_fused` (backward + Adam in one kernel).  Importing this package loads
libfgs_b200.so; there is no fallback if it is missing."""
from .torch_bindings import (FusedAdam, RasterizerSettings, add_noise, diff_rasterize, diff_rasterize_fused,
                             rasterize, relocation_adjustment, update_3d_filter, update_pruning_scores)

__all__ = ['diff_rasterize', 'diff_rasterize_fused', 'rasterize', 'update_pruning_scores', 'RasterizerSettings',
           'FusedAdam', 'update_3d_filter', 'relocation_adjustment', 'add_noise']
