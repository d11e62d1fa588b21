#!/usr/bin/env python
"""Can a pure HBM stream (the dense Adam of sh_rest: 28 B per float, 45 M floats at C2) hide behind the latency-bound front of
the forward pass (preprocess + depth sort + offsets scan + instance generation + tile sort)?  Times, with CUDA events:
the forward alone, the Adam alone, and both launched concurrently (Adam on a side stream), for several grid caps of the
streaming kernel (FGS_ADAM_GRID_CAP is not a product knob: the probe splits the Adam into chunks to throttle it instead).

    python profiles/overlap_probe.py > gpurun_out/r2_overlap_probe.json
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'faster-gaussian-splatting_b200')]


def main():
    import fgs_b200
    from fgs_b200 import _C, _capi, synthetic as syn
    from tests import util
    n, w, h = 1_000_000, 1920, 1080
    sc = syn.make_scene(n, 0).to('cuda')
    cams = [syn.make_camera(k, w, h) for k in range(8)]
    sts = [util.settings_for(c, device='cuda') for c in cams]
    p = torch.zeros(n * 45, device='cuda')
    g, m, v = torch.randn_like(p) * 1e-3, torch.zeros_like(p), torch.zeros_like(p)
    side = torch.cuda.Stream()

    def forward(i):
        return _C.forward(*sc, *sts[i % 8])

    def adam(chunks=1):
        step = n * 45 // chunks
        for c in range(chunks):
            sl = slice(c * step, (c + 1) * step)
            _C.adam_step(g[sl], p[sl], m[sl], v[sl], 1, 1e-4, 0.9, 0.999, 1e-15)

    def timed(fn, reps=30):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def both(i, chunks=1):
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            adam(chunks)
            done = torch.cuda.Event()
            done.record()
        r = forward(i)
        torch.cuda.current_stream().wait_event(done)
        return r

    out = {'forward_alone_ms': timed(forward), 'adam_sh_rest_alone_ms': timed(lambda i: adam())}
    out['serial_ms'] = timed(lambda i: (adam(), forward(i)))
    for chunks in (1, 4, 16):
        out[f'concurrent_ms_chunks{chunks}'] = timed(lambda i: both(i, chunks))
    # per-kernel times of the forward while the stream runs next to it
    _capi.profile_enable(True)
    for i in range(5):
        both(i)
        torch.cuda.synchronize()
        recs = _capi.profile_read()
    _capi.profile_enable(False)
    out['forward_kernels_under_stream_ms'] = {k: round(ms, 4) for k, ms in recs if k != 'adam_step'}
    _capi.profile_enable(True)
    for i in range(5):
        forward(i)
        torch.cuda.synchronize()
        recs = _capi.profile_read()
    _capi.profile_enable(False)
    out['forward_kernels_alone_ms'] = {k: round(ms, 4) for k, ms in recs}
    print(json.dumps({k: (round(x, 4) if isinstance(x, float) else x) for k, x in out.items()}, indent=1))


if __name__ == '__main__':
    main()
