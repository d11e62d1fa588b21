from .rasterization import diff_rasterize, diff_rasterize_fused, rasterize, update_pruning_scores, RasterizerSettings
from .adam import FusedAdam
from .extras import update_3d_filter, relocation_adjustment, add_noise

__all__ = ['diff_rasterize', 'diff_rasterize_fused', 'rasterize', 'update_pruning_scores', 'RasterizerSettings',
           'FusedAdam', 'update_3d_filter', 'relocation_adjustment', 'add_noise']
