"""fgs_b200: Blackwell-native (sm_100a) drop-in for the FasterGSCudaBackend operator API.

    from fgs_b200 import diff_rasterize, rasterize, RasterizerSettings, FusedAdam

The eight public names of the reference backend (/root/reference/FasterGSCudaBackend/__init__.py:14-18)
plus `diff_rasterize_fused` (backward + Adam in one kernel).  Accessing any of them loads
libfgs_b200.so; there is no fallback if it is missing (ImportError).  The helper modules
`fgs_b200.synthetic` (benchmark scenes) and `fgs_b200.parallel` (view-parallel all-reduce) are pure
Python and do not load the library.  Callers either side of the rasterizer: `fgs_b200.loss`
(L1 + DSSIM fused with its gradient) and `fgs_b200.densification` (fused prune / densify / reorder of the
parameters and their Adam state)."""
import importlib

__all__ = ['diff_rasterize', 'diff_rasterize_fused', 'rasterize', 'update_pruning_scores', 'RasterizerSettings',
           'FusedAdam', 'update_3d_filter', 'relocation_adjustment', 'add_noise']


def __getattr__(name):
    if name in __all__:
        tb = importlib.import_module('.torch_bindings', __name__)
        return getattr(tb, name)
    if name in ('_C', '_capi', 'torch_bindings', 'synthetic', 'parallel', 'loss', 'densification', 'fourd'):
        return importlib.import_module('.' + name, __name__)
    if name == 'join_deferred':  # deferred SH schedule (diff_rasterize_fused(..., defer_sh=True)): see _C.join_deferred
        return importlib.import_module('._C', __name__).join_deferred
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')
