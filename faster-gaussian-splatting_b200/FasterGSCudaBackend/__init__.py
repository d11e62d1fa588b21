"""Drop-in alias: `import FasterGSCudaBackend` resolves to the B200-native backend.

Put `faster-gaussian-splatting_b200/` on sys.path (or copy this directory next to Renderer.py as
`Methods/FasterGS/FasterGSCudaBackend`) and the reference's Renderer.py / Model.py / Trainer.py import
the same eight names they import from the original extension (Renderer.py:16, Model.py:17)."""
from fgs_b200 import (FusedAdam, RasterizerSettings, add_noise, diff_rasterize, diff_rasterize_fused, rasterize,
                      relocation_adjustment, update_3d_filter, update_pruning_scores)
from fgs_b200 import _C  # noqa: F401  (the reference's bindings do `from FasterGSCudaBackend import _C`)

__all__ = ['diff_rasterize', 'rasterize', 'update_pruning_scores', 'RasterizerSettings', 'FusedAdam',
           'update_3d_filter', 'relocation_adjustment', 'add_noise']
