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
    box).  Returns the module `FasterGSCudaBackend.torch_bindings` or None."""
    so = list((REF_DIR / 'FasterGSCudaBackend').glob('_C*.so')) if REF_DIR.exists() else []
    if not so:
        return None
    mod = sys.modules.get('FasterGSCudaBackend')
    if mod is not None and str(REF_DIR) not in str(getattr(mod, '__path__', [''])[0]):
        for k in [k for k in sys.modules if k == 'FasterGSCudaBackend' or k.startswith('FasterGSCudaBackend.')]:
            del sys.modules[k]
    sys.path.insert(0, str(REF_DIR))
    try:
        return importlib.import_module('FasterGSCudaBackend.torch_bindings')
    finally:
        sys.path.remove(str(REF_DIR))


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
