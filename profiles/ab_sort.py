"""A/B micro-benchmark of the radix sort alone.

    python profiles/ab_sort.py [lib.so ...]      (default: the product library)

Every library variant is loaded side by side (ctypes.CDLL per path); for each workload the median of
12 timed sorts (CUDA events on the launching stream) is printed.  Workloads mirror C2: the depth sort
(614 k visible of 1 M, 32-bit float keys) and the tile sort (2.15 M instances, 14-bit keys, 16-bit type).
"""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
P = C.c_void_p
ARGS = [P, P, P, P, P, C.c_int32, C.c_int32, P, C.c_size_t, C.POINTER(C.c_int32), P]


def load(path):
    lib = C.CDLL(str(path))
    for name in ('fgs_sort_pairs_u32', 'fgs_sort_pairs_u16'):
        getattr(lib, name).restype = C.c_int
        getattr(lib, name).argtypes = ARGS
    lib.fgs_sort_scratch_bytes.restype = C.c_size_t
    lib.fgs_sort_scratch_bytes.argtypes = [C.c_int32, C.c_int32]
    return lib


WORKLOADS = [  # (n_max, n, end_bit, key bits, label)
    (1_000_000, 614_000, 32, 32, 'depth 614k/1M f32 keys'),
    (2_150_000, 2_150_000, 14, 16, 'tile 2.15M 14-bit u16'),
    (6_000_000, 6_000_000, 14, 16, 'tile 6M 14-bit u16'),
    (2_150_000, 2_150_000, 14, 32, 'tile 2.15M 14-bit u32'),
]


def run(lib, n_max, n, end_bit, kbits):
    g = torch.Generator(device='cuda').manual_seed(0)
    if end_bit == 32:  # depth-like keys: float bits of positive numbers
        keys0 = (torch.rand(n_max, device='cuda', generator=g) * 8 + 0.2).view(torch.int32)
    else:
        keys0 = torch.randint(0, 1 << end_bit, (n_max,), dtype=torch.int32, device='cuda', generator=g)
        if kbits == 16:
            keys0 = keys0.to(torch.int16)
    vals0 = torch.arange(n_max, dtype=torch.int32, device='cuda')
    n_dev = torch.tensor([n], dtype=torch.int32, device='cuda')
    sb = lib.fgs_sort_scratch_bytes(n_max, end_bit)
    scratch = torch.empty(sb, dtype=torch.uint8, device='cuda')
    ka, va = torch.empty_like(keys0), torch.empty_like(vals0)
    fn = lib.fgs_sort_pairs_u32 if kbits == 32 else lib.fgs_sort_pairs_u16
    ts = []
    ok = True
    for it in range(12):
        k, v = keys0.clone(), vals0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        alt = C.c_int32(0)
        torch.cuda.synchronize()
        e0.record()
        rc = fn(k.data_ptr(), ka.data_ptr(), v.data_ptr(), va.data_ptr(), n_dev.data_ptr(), n_max, end_bit,
                scratch.data_ptr(), sb, C.byref(alt), torch.cuda.current_stream().cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        ts.append(e0.elapsed_time(e1))
        if it == 0:  # correctness of the variant
            ks = (ka if alt.value else k)[:n]
            vs = (va if alt.value else v)[:n]
            ref = keys0[:n].to(torch.int64) & (0xffffffff if kbits == 32 else 0xffff)
            order = torch.sort(ref, stable=True).indices
            ok = bool(torch.equal(vs, vals0[:n][order]) and torch.equal(ks, keys0[:n][order]))
    ts.sort()
    return ts[len(ts) // 2] * 1000, ok


def main():
    paths = sys.argv[1:] or [ROOT / 'faster-gaussian-splatting_b200' / 'fgs_b200' / 'lib' / 'libfgs_b200.so']
    libs = [(Path(p).stem, load(p)) for p in paths]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for n_max, n, end_bit, kbits, label in WORKLOADS:
        k = torch.randint(0, 1 << min(end_bit, 31), (n,), dtype=torch.int32, device='cuda')
        torch.sort(k)
        e0.record(); torch.sort(k); e1.record(); torch.cuda.synchronize()
        print(f'## {label}   (torch.sort keys only: {e0.elapsed_time(e1) * 1000:.1f} us)')
        for name, lib in libs:
            us, ok = run(lib, n_max, n, end_bit, kbits)
            print(f'   {name:28s} {us:8.1f} us   {"ok" if ok else "WRONG RESULT"}', flush=True)


if __name__ == '__main__':
    main()
