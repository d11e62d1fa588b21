"""Operators of the reference API that are outside this round's hot path.

Same names and argument lists as /root/reference/FasterGSCudaBackend/FasterGSCudaBackend/
torch_bindings/filter3d.py:6-35 and densification.py:6-22; they raise NotImplementedError (both are
disabled in the benchmarked configuration: FILTER_3D.USE false, USE_MCMC false)."""
from .. import _C


def update_3d_filter(*args, **kwargs):
    return _C.update_3d_filter(*args, **kwargs)


def relocation_adjustment(*args, **kwargs):
    return _C.relocation_adjustment(*args, **kwargs)


def add_noise(*args, **kwargs):
    return _C.add_noise(*args, **kwargs)
