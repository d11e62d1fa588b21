"""Shared helpers for the tests: scenes, settings, access to the compiled reference."""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF_DIR = ROOT / 'baseline' / '_ref'


def settings_for(cam, active_sh_bases=16, bg=(0.0, 0.0, 0.0), aa=False, device='cpu'):
    """13-tuple in RasterizerSettings order."""
    return (cam.w2c.to(device), cam.position.to(device), torch.tensor(bg, dtype=torch.float32, device=device),
            active_sh_bases, cam.width, cam.height, cam.focal_x, cam.focal_y, cam.center_x, cam.center_y,
            cam.near_plane, cam.far_plane, aa)


def load_reference():
    """The UNMODIFIED reference extension built into baseline/_ref (git-ignored, travels to the GPU
    box).  Returns the module `FasterGSCudaBackend.torch_bindings` of the REFERENCE or None.

    The pip-installed reference has no FasterGSCudaBackend/__init__.py (namespace package), so a
    plain import would lose against this repo's drop-in alias package of the same name; the package
    object is therefore created by hand with __path__ pinned to baseline/_ref."""
    ref_pkg = REF_DIR / 'FasterGSCudaBackend'
    if not (ref_pkg.exists() and list(ref_pkg.glob('_C*.so'))):
        return None
    import types
    import torch  # noqa: F401  (libc10.so etc. must be loaded before the extension)
    mod = sys.modules.get('FasterGSCudaBackend')
    if mod is None or [str(p) for p in getattr(mod, '__path__', [])] != [str(ref_pkg)]:
        for k in [k for k in sys.modules if k == 'FasterGSCudaBackend' or k.startswith('FasterGSCudaBackend.')]:
            del sys.modules[k]
        pkg = types.ModuleType('FasterGSCudaBackend')
        pkg.__path__ = [str(ref_pkg)]
        sys.modules['FasterGSCudaBackend'] = pkg
    tb = importlib.import_module('FasterGSCudaBackend.torch_bindings')
    assert str(ref_pkg) in tb.__file__, tb.__file__
    c_mod = sys.modules['FasterGSCudaBackend._C']
    assert str(ref_pkg) in c_mod.__file__ and c_mod.__file__.endswith('.so'), c_mod.__file__
    return tb


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """|a - b|_1 / |b|_1 (scale-free aggregate error; per-element rtol is meaningless for gradients
    that are sums of many float atomics of mixed sign)."""
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).abs().sum() / b.abs().sum().clamp_min(1e-30))


def cos_sim(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    if float(a.abs().sum()) == 0.0 and float(b.abs().sum()) == 0.0:
        return 1.0
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))
