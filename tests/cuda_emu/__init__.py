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


# the rasterizer proper: c_abi.cu defines fgs::report_error / fgs_last_error itself
FULL = ['preprocess', 'sort', 'binning', 'blend_forward', 'blend_backward', 'preprocess_backward', 'aux_ops', 'c_abi']


def _compile(src: Path, sanitize_threads: bool, deps_mtime: float) -> str:
    """One object per source file, shared by every library that contains it."""
    obj = BUILD / f"{src.stem}{'_tsan' if sanitize_threads else ''}.o"
    if not obj.exists() or obj.stat().st_mtime < max(deps_mtime, src.stat().st_mtime):
        cmd = ['g++', '-std=c++20', '-O1', '-g', '-pthread', '-fPIC', '-DFGS_CUDA_EMU=1', '-Wno-unknown-pragmas', '-Wno-attributes',
               '-include', str(HERE / 'cuda_emu.h'), '-x', 'c++', '-c', str(src), '-o', str(obj)]
        if sanitize_threads:
            cmd[1:1] = ['-fsanitize=thread']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'emulator build of {src.name} failed:\n{r.stderr[-6000:]}')
    return str(obj)


def build(name: str, sanitize_threads: bool = False) -> ctypes.CDLL:
    """`name`: one self-contained operator file (loss, densify, peer_adam), 'full' = the whole rasterizer, or 'all' =
    every symbol of include/fgs_b200.h in one library (what FGS_B200_LIB can point the Python layer at)."""
    from concurrent.futures import ThreadPoolExecutor
    names = FULL if name == 'full' else FULL + ['loss', 'densify', 'peer_adam', 'slice4d'] if name == 'all' else [name]
    srcs = [CSRC / f'{n}.cu' for n in names] + ([] if name in ('full', 'all') else [HERE / 'emu_support.cpp'])
    tag = f'{name}_tsan' if sanitize_threads else name
    out = BUILD / f'libfgs_emu_{tag}.so'
    common = [HERE / 'cuda_emu.h', CSRC / 'fgs_launch.h', CSRC / 'fgs_device.cuh', CSRC / 'fgs_host.h', ROOT / 'include' / 'fgs_b200.h',
              Path(__file__)]
    deps_mtime = max(p.stat().st_mtime for p in common)
    if not out.exists() or out.stat().st_mtime < max(deps_mtime, max(p.stat().st_mtime for p in srcs)):
        BUILD.mkdir(exist_ok=True)
        with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
            objs = list(ex.map(lambda s: _compile(s, sanitize_threads, deps_mtime), srcs))
        cmd = ['g++', '-shared', '-pthread', '-o', str(out), *objs]
        if sanitize_threads:
            cmd[1:1] = ['-fsanitize=thread']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'emulator link failed:\n{r.stderr[-4000:]}')
    return ctypes.CDLL(str(out))


def library_path(name: str) -> Path:
    build(name)
    return BUILD / f'libfgs_emu_{name}.so'
