"""Pinned host -> device copy bandwidth on this box (what bounds bench.py's e2e input pipeline)."""
import torch
dev = torch.device('cuda')
for mb in (6, 25, 100):
    n = mb * 1024 * 1024
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ('one stream', 'two streams (halves)'):
        ts = []
        for _ in range(6):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if mode == 'one stream':
                dst.copy_(src, non_blocking=True)
            else:
                for st, sl in ((s1, slice(0, n // 2)), (s2, slice(n // 2, n))):
                    st.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(st):
                        dst[sl].copy_(src[sl], non_blocking=True)
                    torch.cuda.current_stream().wait_stream(st)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f'{mb:4d} MB {mode:22s}: {t:.3f} ms  {n / t / 1e6:.1f} GB/s')
