#!/usr/bin/env python
"""One launch of EVERY kernel of the library at benchmark size, inside a cudaProfilerStart/Stop window, for
`ncu --set full --profile-from-start off` ("every kernel is evidenced by an ncu capture").

    ncu --set full --clock-control none --profile-from-start off -k regex:'<see profiles/r2_call14_ncu.sh>' -o /tmp/r2_new_kernels python profiles/evidence_run.py
    ncu -i /tmp/r2_new_kernels.ncu-rep --page raw --csv > gpurun_out/r2_new_kernels_raw.csv     (on the box: the report is too big to travel)
    python profiles/summarize_ncu_r2.py gpurun_out/r2_new_kernels_raw.csv > profiles/r2_kernels_ncu.md
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'faster-gaussian-splatting_b200'), str(ROOT / 'profiles')]


def main():
    import fgs_b200
    from fgs_b200 import densification as D, fourd as F4, loss as L, synthetic as syn
    from tests import util
    from tests.test_widen_fourd import make_scene_4d
    import bench_widen
    dev = torch.device('cuda', 0)
    n, w, h = 1_000_000, 1920, 1080
    sc = syn.make_scene(n, 0).to(dev)
    lrs = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]
    rs = [fgs_b200.RasterizerSettings(*util.settings_for(syn.make_camera(k, w, h), device=dev)) for k in range(4)]
    gt8 = (syn.make_target(w, h) * 255).round().to(torch.uint8).to(dev)
    params = [p.clone().requires_grad_(True) for p in sc]
    opt = fgs_b200.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs)], lr=0.0, eps=1e-15)
    dens = torch.zeros(2, n, device=dev)

    def step(k, mode):
        if mode == 'unfused':
            img = fgs_b200.diff_rasterize(*params, dens, rs[k])
        else:
            img = fgs_b200.diff_rasterize_fused(*params, dens, rs[k], opt, defer_sh=(mode == 'deferred'))
        terms, g = L.l1_dssim_loss_and_grad(img.detach(), gt8, 1.0, 0.0)
        img.backward(g)
        opt.step()
        opt.zero_grad()

    p4 = [p.to(dev).requires_grad_(True) for p in make_scene_4d(500_000, 0, w)]

    def step4d():
        img = F4.diff_rasterize_4d(*p4, 0.37, torch.empty(0), rs[1])
        terms, g = L.l1_dssim_loss_and_grad(img.detach(), gt8, 1.0, 0.0)
        img.backward(g)
        for p in p4:
            p.grad = None

    prof = {'on': False}

    class window:  # cudaProfilerStart/Stop around the calls whose kernels are wanted (ncu --profile-from-start off)
        def __enter__(self):
            if prof['on']:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()

        def __exit__(self, *exc):
            if prof['on']:
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()
            return False

    def everything():
        # (the ten kernels of a plain training step -- preprocess, onesweep x2 types, offsets_scan, create_instances x2,
        # ranges_gather, blend forward / backward, preprocess backward -- are in profiles/r1_kernels_ncu.md; the windows below
        # hold everything else)
        img = fgs_b200.diff_rasterize(*params, dens, rs[0])
        with window():
            terms, g = L.l1_dssim_loss_and_grad(img.detach(), gt8, 1.0, 0.0)   # l1_kernel
        img.backward(g)
        with window():
            opt.step()                                                          # adam_kernel x6
        opt.zero_grad()
        step(1, 'fused')
        step(2, 'deferred')
        with window():
            step(3, 'deferred')  # preprocess (geometry only), color_kernel, preprocess_backward (deferred), adam_rest_kernel
            fgs_b200.join_deferred()
        with torch.no_grad():
            det = [p.detach() for p in params]
            with window():
                fgs_b200.rasterize(*det, rs[0], to_chw=True)            # inference instantiations
                fgs_b200.rasterize(*det, rs[0], to_chw=False, clamp_output=False)
                scores = torch.zeros(n, device=dev)
                fgs_b200.update_pruning_scores(scores, *det, rs[0])       # blend_backward_kernel<true>
            img = fgs_b200.rasterize(*det, rs[0], to_chw=True)
        with window():
            L.l1_dssim_loss_and_grad(img, gt8, 0.8, 0.2)                    # ssim_stats + ssim_grad
            x = img.clone().requires_grad_(True)
            L.l1_dssim_loss(x, gt8, 0.8, 0.2).backward()                   # the autograd form (forward, then backward kernel)
            step4d()                                                        # slice4d fwd/bwd, covariance-input preprocess / backward
            with torch.no_grad():
                F4.rasterize_4d(*[p.detach() for p in p4], 0.6, rs[2], to_chw=True)
        # densification / reordering on a 1 M optimizer state
        o = bench_widen.make_optimizer(n, dev)
        g = torch.Generator().manual_seed(1)
        info = torch.stack([torch.randint(0, 40, (n,), generator=g).float(), torch.rand(n, generator=g) * 0.012]).to(dev)
        with window():
            D.adaptive_density_control(o, info, 0.0002, 0.005, True, 0.01, 4.0)
            means = o.param_groups[0]['params'][0].detach()
            D.sort_param_groups(o, D.morton_order(means))
        keep = (torch.rand(means.shape[0], generator=g) < 0.7).to(dev)
        with window():
            D.prune_param_groups(o, keep)
            D.reset_state(o, torch.arange(0, 1000, device=dev))
            # the small operators around the rasterizer
            from tests import test_aux_ops as TA
            d, filt = TA._inputs('cuda')
            TA._run(fgs_b200, d, filt, fgs_b200._C.add_noise)
        del o
        torch.cuda.synchronize()

    everything()  # warm-up: allocator, capacity hints, function attributes
    torch.cuda.synchronize()
    prof['on'] = True
    everything()
    print('evidence run done')


if __name__ == '__main__':
    main()
