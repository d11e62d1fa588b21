"""Build libfgs_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python faster-gaussian-splatting_b200/build.py [--force] [--verbose]

No torch involved: the library only depends on the (statically linked) CUDA runtime, so it can be
dlopen'ed on a machine without a GPU (the driver is only needed at the first CUDA call).
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / 'csrc'
OUT_DIR = HERE / 'fgs_b200' / 'lib'
LIB = OUT_DIR / 'libfgs_b200.so'
OBJ_DIR = HERE / 'build' / 'obj'

SOURCES = ['preprocess.cu', 'sort.cu', 'binning.cu', 'blend_forward.cu', 'blend_backward.cu',
           'preprocess_backward.cu', 'aux_ops.cu', 'peer_adam.cu', 'loss.cu', 'densify.cu', 'slice4d.cu', 'c_abi.cu']
HEADERS = ['fgs_device.cuh', 'fgs_host.h', 'fgs_launch.h', '../../include/fgs_b200.h']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo', '--use_fast_math',
    '-Xcompiler', '-fPIC,-O3',
    '-Xptxas', '-v',
]


def find_nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not Path(nvcc).exists():
        raise RuntimeError('nvcc not found; libfgs_b200.so cannot be built')
    return nvcc


def newest_mtime(paths) -> float:
    return max(p.stat().st_mtime for p in paths)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: Path | None = None) -> Path:
    """`defines`/`out` build an experiment variant (e.g. -DFGS_UNROLL_BWD=1) next to the product library;
    select it at run time with FGS_B200_LIB=<path>."""
    global LIB, OBJ_DIR
    if out is not None:
        LIB = Path(out)
        OBJ_DIR = HERE / 'build' / ('obj_' + LIB.stem)
        force = True
    srcs = [CSRC / s for s in SOURCES]
    deps = srcs + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    if LIB.exists() and not force and LIB.stat().st_mtime >= newest_mtime(deps):
        return LIB
    nvcc = find_nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    OUT_DIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: Path) -> tuple[Path, str]:
        obj = OBJ_DIR / (src.stem + '.o')
        cmd = [nvcc, *NVCC_FLAGS, *[f'-D{d}' for d in defines], '-c', str(src), '-o', str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}')
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, srcs))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    (HERE / 'build' / 'ptxas.log').write_text('\n'.join(log for _, log in results))
    objs = [str(o) for o, _ in results]
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC',
           '-cudart', 'static', '-o', str(LIB), *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    ap.add_argument('-D', dest='defines', action='append', default=[])
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    print(build(force=args.force, verbose=args.verbose, defines=tuple(args.defines), out=args.out))
