"""TEST INFRASTRUCTURE ONLY: builds csrc/<name>.cu with g++ against cuda_emu.h (a thread-per-CUDA-thread emulator)
into tests/cuda_emu/_build/libfgs_emu_<name>.so, so that the kernel SOURCE and its C-ABI launch logic can be executed
on host memory in a container without a GPU.  The product never imports this package."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / 'faster-gaussian-splatting_b200' / 'csrc'
BUILD = HERE / '_build'


def build(name: str, sanitize_threads: bool = False) -> ctypes.CDLL:
    src = CSRC / f'{name}.cu'
    tag = f'{name}_tsan' if sanitize_threads else name
    out = BUILD / f'libfgs_emu_{tag}.so'
    deps = [src, HERE / 'cuda_emu.h', HERE / 'emu_support.cpp', CSRC / 'fgs_launch.h', ROOT / 'include' / 'fgs_b200.h',
            Path(__file__)]
    if not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        BUILD.mkdir(exist_ok=True)
        cmd = ['g++', '-std=c++20', '-O1', '-g', '-pthread', '-shared', '-fPIC', '-DFGS_CUDA_EMU=1',
               '-Wno-unknown-pragmas', '-Wno-attributes', '-include', str(HERE / 'cuda_emu.h'),
               '-x', 'c++', str(src), str(HERE / 'emu_support.cpp'), '-o', str(out)]
        if sanitize_threads:
            cmd[1:1] = ['-fsanitize=thread']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'emulator build of {src.name} failed:\n{r.stderr[-4000:]}')
    return ctypes.CDLL(str(out))
