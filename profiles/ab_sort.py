"""A/B micro-benchmark of the radix sort alone (FGS_B200_LIB selects the variant)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / 'faster-gaussian-splatting_b200'))
from fgs_b200 import _capi

lib = _capi.load()
for n, end_bit in [(614_000, 32), (2_150_000, 14), (6_000_000, 14)]:
    hi = 1 << min(end_bit, 31)
    keys0 = torch.randint(0, hi, (n,), dtype=torch.int32, device='cuda')
    if end_bit == 32:  # depth-like keys: float bits of positive numbers
        keys0 = (torch.rand(n, device='cuda') * 8 + 0.2).view(torch.int32)
    vals0 = torch.arange(n, dtype=torch.int32, device='cuda')
    n_dev = torch.tensor([n], dtype=torch.int32, device='cuda')
    sb = lib.fgs_sort_scratch_bytes(n, end_bit)
    scratch = torch.empty(sb, dtype=torch.uint8, device='cuda')
    ka, va = torch.empty_like(keys0), torch.empty_like(vals0)
    ts = []
    for it in range(12):
        k, v = keys0.clone(), vals0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        alt = C.c_int32(0)
        e0.record()
        lib.fgs_sort_pairs_u32(k.data_ptr(), ka.data_ptr(), v.data_ptr(), va.data_ptr(), n_dev.data_ptr(), n, end_bit,
                               scratch.data_ptr(), sb, C.byref(alt), torch.cuda.current_stream().cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = keys0.clone()
    e0.record(); torch.sort(k); e1.record(); torch.cuda.synchronize()
    print(f'n={n} bits={end_bit}: median {ts[len(ts)//2]*1000:.1f} us  (torch.sort {e0.elapsed_time(e1)*1000:.1f} us)')
