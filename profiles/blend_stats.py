#!/usr/bin/env python
"""Issue roofline of the blend kernels: counters from the instrumentation build of the library
(-DFGS_BLEND_STATS=1: pairs executed by the inner loops, lanes doing useful work, flush statistics) combined with the
warp-instruction counts of the committed ncu capture.

    python faster-gaussian-splatting_b200/build.py -D FGS_BLEND_STATS=1 --out $PWD/faster-gaussian-splatting_b200/fgs_b200/lib/libfgs_b200_stats.so
    FGS_B200_LIB=$PWD/faster-gaussian-splatting_b200/fgs_b200/lib/libfgs_b200_stats.so python profiles/blend_stats.py > gpurun_out/r2_blend_stats.json
"""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'faster-gaussian-splatting_b200')]


def main():
    import fgs_b200
    from fgs_b200 import _capi, synthetic as syn
    from tests import util
    lib = _capi.load()
    n, w, h = 1_000_000, 1920, 1080
    sc = syn.make_scene(n, 0).to('cuda')
    out = {}
    for view in (0, 3):
        cam = syn.make_camera(view, w, h)
        st = util.settings_for(cam, device='cuda')
        gt = syn.make_target(w, h, seed=view).cuda()
        params = [p.clone().requires_grad_(True) for p in sc]
        buf = (C.c_uint64 * 16)()
        _capi.check(lib.fgs_debug_blend_stats(buf, 1), 'blend_stats')
        img = fgs_b200.diff_rasterize(*params, torch.empty(0), fgs_b200.RasterizerSettings(*st))
        (img - gt).abs().mean().backward()
        _capi.check(lib.fgs_debug_blend_stats(buf, 1), 'blend_stats')
        v = [int(x) for x in buf]
        sm, ghz = 148, 1.965
        issue_rate = sm * 4 * ghz * 1e9  # warp instructions per second at one per scheduler and cycle
        fwd_instr, bwd_instr = 217.7e6, 417.3e6  # smsp__inst_executed.sum, profiles/r1_kernels_ncu.md (same scene, view 0)
        out[f'view{view}'] = {
            'forward': {'pairs': v[0], 'lanes_blended': v[1], 'lanes_live': v[8], 'batches': v[2], 'batches_without_hit': v[3],
                        'useful_lane_fraction': round(v[1] / max(32 * v[0], 1), 4),
                        'live_lane_fraction': round(v[8] / max(32 * v[0], 1), 4),
                        'warp_instr_per_pair': round(fwd_instr / max(v[0], 1), 1),
                        'warp_instr_per_useful_pixel_gaussian': round(fwd_instr / max(v[1], 1), 2),
                        'issue_bound_ms': round(fwd_instr / issue_rate * 1e3, 4)},
            'backward': {'pairs': v[4], 'lanes_contributing': v[5], 'flushes': v[6], 'slots_flushed': v[7],
                         'useful_lane_fraction': round(v[5] / max(32 * v[4], 1), 4),
                         'slots_per_flush': round(v[7] / max(v[6], 1), 2),
                         'warp_instr_per_pair': round(bwd_instr / max(v[4], 1), 1),
                         'warp_instr_per_useful_pixel_gaussian': round(bwd_instr / max(v[5], 1), 2),
                         'issue_bound_ms': round(bwd_instr / issue_rate * 1e3, 4)},
        }
    out['note'] = ('issue_bound_ms = warp instructions / (148 SM x 4 schedulers x 1.965 GHz): the time the kernel would take at '
                   'one instruction per scheduler and cycle; instruction counts are those of view 0 (ncu)')
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
