#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: train iters/s @ 1M Gaussians, 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mine|reference] [--config c2|c3|c5] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one training iteration of the reference's Trainer loop (Trainer.py:170-201): forward + loss + backward +
Adam per view, through the public Python operator API (the call a user of the reference makes).  Synthetic scene
of the named shape (fgs_b200/synthetic.py), random-init parameters, fp32 arithmetic.

Configs (BASELINE.json):  c2 (default) = configs[1], 1 M Gaussians at 1920x1080;  c3 = configs[2], 3 M Gaussians at
1600x1063, FasterGSFused path;  c5 = configs[4], 500 k 4D Gaussians x 150 time steps at 1080p (no reference source).
configs[3] (8-view batches on 1/2/4/8 GPUs) = `--batch 8` with `--gpus N`: the 8 views of a batch are split over the
ranks (strong scaling); without --batch every rank trains one view per step (weak scaling, batch = N).

Prints ONE JSON line on rank 0.  Keys: see DESIGN.md "Measurement".  With N > 1 every rank renders its own views,
the gradients are summed over the ranks and every rank ends the step with identical parameters (--dp-mode selects how:
one kernel over NVLink peer / multicast memory, NCCL reduce-scatter + sharded Adam + all-gather, or NCCL all-reduce +
dense Adam); value = views trained per second over all ranks.  `--impl reference` runs the UNMODIFIED reference CUDA
backend from baseline/_ref through the same loop (north_star: "next to the reference's own CUDA path"); if it is not
installed, the CPU oracle port is timed instead and the line says so.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / 'faster-gaussian-splatting_b200'))
sys.path.insert(0, str(ROOT))

SH_BASES, N_VIEWS = 16, 8
LRS = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]  # means, scales, rotations, opacities, sh0, sh_rest (Trainer.py:56-66)
CONFIGS = {
    'c2': dict(n=1_000_000, w=1920, h=1080, metric='train_iters_per_sec_1M_gaussians_1080p',
               workload='configs[1]: fastergs_garden shape, 1M Gaussians, 1920x1080, 16 SH bases'),
    'c3': dict(n=3_000_000, w=1600, h=1063, metric='train_iters_per_sec_3M_gaussians_1600p',
               workload='configs[2]: bicycle shape, 3M Gaussians, 1600x1063, 16 SH bases, FasterGSFused backward+Adam'),
    'c5': dict(n=500_000, w=1920, h=1080, metric='train_iters_per_sec_500k_4d_gaussians_1080p',
               workload='configs[4]: FasterGS4D shape, 500k 4D Gaussians x 150 time steps, 1920x1080, 16 SH bases'),
}
DP_MODES = ('auto', 'allreduce', 'flat', 'pipelined', 'sharded', 'peer', 'peer-mc', 'peer-rows', 'peer-rows-sync', 'peer-rows-mono', 'peer-rows-mono-p2p')


def peak_hbm_gbs():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        try:
            return float(json.loads(f.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def algorithmic_bytes(N, V, I, P, T, S=45):
    """compulsory bytes per stage and per training iteration."""
    st = {
        'preprocess': 44 * N + (12 + 4 * S) * V + 4 * N + 56 * V,
        'depth sort': 16 * V,
        'create_instances': 48 * V + 6 * I,
        'tile sort': 12 * I,
        'ranges_and_gather': 2 * I + 8 * T,
        'blend_forward': 4 * I + 44 * V + 8 * T + 20 * P,
        'blend_backward': 4 * I + 44 * V + 32 * P + 36 * V,
        'preprocess_backward': 4 * N + (80 + 4 * S) * V + (56 + 4 * S) * V + 16 * V + 236 * N,
        'adam_step': 1652 * N,
    }
    st['iteration'] = 1940 * N + (408 + 12 * S) * V + 28 * I + 52 * P + 16 * T
    st['iteration_fused'] = 1468 * N + (352 + 8 * S) * V + 28 * I + 52 * P + 16 * T
    st['preprocess_backward_fused'] = st['iteration_fused'] - (st['iteration'] - st['preprocess_backward'] - st['adam_step'])
    return st


def make_problem(cfg, views=N_VIEWS):
    from fgs_b200 import synthetic as syn
    n, w, h = cfg['n'], cfg['w'], cfg['h']
    sc = syn.make_scene(n, 0, width=w)
    cams = [syn.make_camera(k, w, h, n_views=views) for k in range(views)]
    # ground-truth images are 8-bit (as image datasets are): the e2e path ships them over PCIe as uint8 and
    # converts on the device; the resident path trains on the very same (quantised) float targets
    targets8 = [(syn.make_target(w, h, seed=k) * 255.0).round().clamp(0, 255).to(torch.uint8) for k in range(views)]
    return sc, cams, targets8


def settings_tuple(cam, device, bg):
    return (cam.w2c.to(device), cam.position.to(device), bg, SH_BASES, cam.width, cam.height, cam.focal_x,
            cam.focal_y, cam.center_x, cam.center_y, cam.near_plane, cam.far_plane, False)


def load_backend(impl):
    if impl == 'reference':
        from tests import util
        return util.load_reference()
    import fgs_b200
    return fgs_b200


_WINDOW = {}


def torch_dssim(x, y):
    """1 - mean SSIM with the 11x11 Gaussian window (sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2): plain torch ops
    for the arms that do not use the fused loss kernel."""
    key = (x.device, x.dtype)
    if key not in _WINDOW:
        g = torch.exp(-((torch.arange(11, dtype=x.dtype, device=x.device) - 5.0) ** 2) / (2 * 1.5 ** 2))
        g = g / g.sum()
        _WINDOW[key] = (g[:, None] * g[None, :]).expand(3, 1, 11, 11).contiguous()
    w = _WINDOW[key]
    conv = lambda t: torch.nn.functional.conv2d(t[None], w, padding=5, groups=3)[0]
    mx, my = conv(x), conv(y)
    sxx, syy, sxy = conv(x * x) - mx * mx, conv(y * y) - my * my, conv(x * y) - mx * my
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    ssim = ((2 * mx * my + c1) * (2 * sxy + c2)) / ((mx * mx + my * my + c1) * (sxx + syy + c2))
    return 1.0 - ssim.mean()


def run_gpu(args, impl):
    cfg = CONFIGS[args.config]
    n_gauss, width, height = cfg['n'], cfg['w'], cfg['h']
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    n_ranks = world  # ranks of the job as launched (the reference arm folds them into one process below)
    batch = args.batch if args.batch else n_ranks  # views per optimizer step over the whole job
    if batch % n_ranks:
        raise SystemExit(f'bench.py: --batch {batch} is not a multiple of the {n_ranks} ranks')
    if impl == 'reference' and world > 1:
        # the reference has no multi-GPU mode (Renderer.py:58-61): its semantics for a B-view batch is
        # gradient accumulation on one GPU; rank 0 runs that alone, the other ranks leave
        if rank != 0:
            return None
        world = 1
    vpr = batch // world  # views per rank and step
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group('nccl', device_id=dev)
    be = load_backend(impl)
    from fgs_b200 import parallel
    if be is None:
        if rank == 0:
            return run_cpu_port(args, reason='baseline/_ref not installed on this box')
        return None

    sc, cams, targets = make_problem(cfg)
    mine = impl == 'mine'
    # ---- how the step ends: who sums the gradients and who applies Adam ------------------------------------------
    dp_mode = args.dp_mode if (dist is not None and mine) else 'none'
    if dp_mode == 'auto':
        # one view per rank and step: compact gradient exchange + sh_rest broadcast under the next forward; several
        # views per rank (gradient accumulation): the plain one-kernel reduce + Adam + broadcast
        dp_mode = 'peer-rows' if vpr == 1 else 'peer'
    opt_obj = None  # PeerAdam / ShardedAdam own flat parameter + gradient buffers
    if dp_mode.startswith('peer'):
        if dp_mode.startswith('peer-rows') and vpr != 1:
            raise SystemExit('bench.py: --dp-mode peer-rows needs one view per rank and step')
        try:
            compact = dp_mode.startswith('peer-rows')
            mc = (dp_mode == 'peer-mc') if not compact else ('auto' if not dp_mode.endswith('-p2p') else False)
            opt_obj = parallel.PeerAdam([p.to(dev).contiguous() for p in sc], LRS, multicast=mc, compact=compact,
                                        active_sh_bases=SH_BASES, overlap_sh_broadcast=('sync' not in dp_mode),
                                        split=('mono' not in dp_mode),
                                        small_multicast={'auto': 'auto', 'on': True, 'off': False}[args.small_mc])
            if args.side_ctas:
                opt_obj.side_ctas_per_sm = args.side_ctas
        except Exception as e:  # no symmetric memory / no NVLS on this box: NCCL collectives instead, and say so
            if args.dp_mode != 'auto':
                raise
            print(f'[bench] PeerAdam unavailable ({e!r}); falling back to --dp-mode sharded', file=sys.stderr)
            dp_mode = 'sharded'
    if dp_mode == 'sharded':
        opt_obj = parallel.ShardedAdam([p.to(dev).contiguous() for p in sc], LRS)
    if opt_obj is not None:
        params, opt = opt_obj.params, opt_obj
    else:
        params = [p.to(dev).contiguous().requires_grad_(True) for p in sc]
        opt = be.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(params, LRS)], lr=0.0, eps=1e-15)
    flat = parallel.FlatGradients(params) if dp_mode == 'flat' else None
    registered = opt_obj if opt_obj is not None else flat  # object whose flat gradient buffer the backward writes into
    bg = torch.zeros(3, device=dev)
    st_dev = [settings_tuple(c, dev, bg) for c in cams]
    gt_dev = [t.to(dev).float().div_(255.0) for t in targets]
    gt8_dev = [t.to(dev) for t in targets]
    gt_pin = [t.pin_memory() for t in targets]
    w2c_pin = [c.w2c.pin_memory() for c in cams]
    pos_pin = [c.position.pin_memory() for c in cams]
    empty = torch.empty(0)

    def view_of(mi):  # micro-step mi of rank r trains view (mi * world + r) of the 8-view ring
        return (mi * world + rank) % N_VIEWS

    # FasterGSFused semantics (backward applies Adam in the same kernel) needs the whole batch in one backward
    use_fused = mine and dist is None and vpr == 1 and not args.unfused
    fused_loss = mine and not args.torch_loss
    # deferred SH schedule: the sh_rest part of the fused step runs on a side stream under the next forward's geometry pass
    defer_sh = use_fused and args.defer_sh  # measured +0.5 % at C2 (profiles/r2_results.md): opt-in
    if mine and args.defer_ctas:
        from fgs_b200 import _capi as _capi0
        _capi0.load().fgs_debug_set_deferred_ctas_per_sm(args.defer_ctas)
    if mine and args.overlap_adam:
        from fgs_b200 import _capi
        _capi.load().fgs_debug_set_overlap_adam(1)
    if mine:
        from fgs_b200 import _C as mine_C, loss as loss_mod
    lam = (0.8, 0.2) if args.yaml_loss else (1.0, 0.0)
    dens_info = torch.zeros(2, n_gauss, device=dev) if args.yaml_loss else empty

    def train_view(st, gt, fused, j=0):
        """forward + loss + backward of one view; gt is uint8 or float on the device."""
        rs = be.RasterizerSettings(*st)
        if registered is not None and vpr > 1:
            # the first view of the step writes its gradients straight into the flat buffer, the following views
            # into fresh tensors that autograd adds onto it
            mine_C.set_gradient_outputs(registered.grads if opt_obj is not None else registered.views) if j == 0 else mine_C.set_gradient_outputs(None)
        if fused:  # backward applies the Adam update in the same kernel; opt.step() is then a no-op
            img = be.diff_rasterize_fused(*params, dens_info, rs, opt, defer_sh=defer_sh)
        else:
            img = be.diff_rasterize(*params, dens_info, rs)
        if fused_loss:  # csrc/loss.cu: loss value and d loss / d image in one kernel, uint8 or float target
            terms, grad_image = loss_mod.l1_dssim_loss_and_grad(img.detach(), gt, lam[0], lam[1])
            img.backward(grad_image)
            return terms[0]
        if gt.dtype == torch.uint8:
            gt = gt.float().div_(255.0)  # uint8 -> float on the device
        if args.yaml_loss:  # torch L1 + a torch conv2d DSSIM (NeRFICG's fused_dssim is not in the reference tree)
            loss = lam[0] * (img - gt).abs().mean() + lam[1] * torch_dssim(img, gt)
        else:
            loss = (img - gt).abs().mean()
        loss.backward()
        return loss

    dp_events = []  # (start, end) CUDA events around the end of the step while profile_mine runs

    def finish_step():
        if dp_events is not None and dp_probe['on']:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            finish_step_inner()
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            dp_events.append((e0, e1))
        else:
            finish_step_inner()

    dp_probe = {'on': False}

    def finish_step_inner():
        if dist is not None and opt_obj is None:
            if dp_mode == 'flat':
                flat.allreduce(params)
            elif dp_mode == 'pipelined':
                parallel.allreduce_and_step(params, opt)
                opt.zero_grad()
                return
            else:
                parallel.allreduce_gradients(params)
        opt.step()
        opt.zero_grad()

    def step_resident(it, fused=None):
        f = use_fused if fused is None else fused
        loss = None
        for j in range(vpr):
            k = view_of(it * vpr + j)
            loss = train_view(st_dev[k], gt8_dev[k] if fused_loss else gt_dev[k], f, j)
        finish_step()
        return loss

    # end to end: host buffers in, loss out.  Every view's target image and camera are copied from pinned
    # host memory (double-buffered on a copy stream so the transfer overlaps the previous view's kernels -- an input
    # pipeline, as in any trainer), the loss is read back every step.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{'gt8': torch.empty_like(gt8_dev[0]), 'w2c': torch.empty(4, 4, device=dev), 'pos': torch.empty(3, device=dev),
             'ready': torch.cuda.Event(), 'free': torch.cuda.Event()} for _ in range(2)]

    def prefetch(mi):
        b = bufs[mi % 2]
        k = view_of(mi)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(b['free'])
            b['gt8'].copy_(gt_pin[k], non_blocking=True)
            b['w2c'].copy_(w2c_pin[k], non_blocking=True)
            b['pos'].copy_(pos_pin[k], non_blocking=True)
            b['ready'].record(copy_stream)

    e2e_state = {'primed': -1}
    # every step's loss is copied to pinned host memory and read on the host inside the timed region,
    # one step late (asynchronous metric read-back, as a trainer's logger does); the last one is
    # drained by the synchronize that closes the timed region
    loss_pin = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    loss_log = []

    def read_loss(it, loss):
        slot = it % 2
        loss = loss.detach().reshape(1)
        loss_ready[slot].record(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):  # device -> host on the copy stream: the compute stream does not wait for it
            copy_stream.wait_event(loss_ready[slot])
            loss_pin[slot].copy_(loss, non_blocking=True)
            loss.record_stream(copy_stream)
            loss_evt[slot].record(copy_stream)
        prev = (it - 1) % 2
        if it > 0 and loss_evt[prev].query() is not None:
            loss_evt[prev].synchronize()
            loss_log.append(float(loss_pin[prev][0]))
        return loss_log[-1] if loss_log else 0.0

    diag = {'h2d': True, 'd2h': True}  # FGS_BENCH_DIAG switches the two halves of the e2e path off in turn

    def step_e2e(it, fused=None):
        f = use_fused if fused is None else fused
        if not diag['h2d']:
            loss = step_resident(it, fused)
            return read_loss(it, loss) if diag['d2h'] else loss
        loss = None
        for j in range(vpr):
            mi = it * vpr + j
            if e2e_state['primed'] < mi:
                for b in bufs:
                    b['free'].record(torch.cuda.current_stream())
                prefetch(mi)
            prefetch(mi + 1)
            e2e_state['primed'] = mi + 1
            b = bufs[mi % 2]
            k = view_of(mi)
            torch.cuda.current_stream().wait_event(b['ready'])
            st = (b['w2c'], b['pos']) + st_dev[k][2:]
            if fused_loss:  # the loss kernel reads the uint8 staging buffer itself: released after the view
                loss = train_view(st, b['gt8'], f, j)
                b['free'].record(torch.cuda.current_stream())
            else:
                gt = b['gt8'].float().div_(255.0)  # uint8 -> float on the device
                b['free'].record(torch.cuda.current_stream())
                loss = train_view(st, gt, f, j)
        finish_step()
        return read_loss(it, loss) if diag['d2h'] else loss

    init_params = [p.detach().clone() for p in params]

    def reset_state():
        """Every timed loop starts from the same scene and a fresh optimizer: the cost of a step drifts as
        training moves the Gaussians, so loops that start from different states are not comparable."""
        if mine:
            mine_C.join_deferred(dev)  # a deferred sh_rest update may still be in flight
        with torch.no_grad():
            for p, p0 in zip(params, init_params):
                p.copy_(p0)
                p.grad = None
        if opt_obj is not None:
            opt_obj.reset()
        else:
            opt.state.clear()
        if dens_info.numel():
            dens_info.zero_()

    def timed(fn, steps, warmup, sample_clocks=False):
        reset_state()
        e2e_state['primed'] = -1
        smp = sampler if (sample_clocks and rank == 0) else None
        if smp:  # sampled from the warm-up on (same load as the timed steps): the timed region alone can be
            smp.start()  # shorter than nvidia-smi's start-up + sampling period
        for it in range(warmup):
            fn(it)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for it in range(steps):
            fn(warmup + it)
        if mine:
            mine_C.join_deferred(dev)  # the timed region ends when the last step's deferred update has finished, too
        e1.record()
        torch.cuda.synchronize()
        clocks = None
        if sample_clocks:
            # keep the identical load running (untimed) until nvidia-smi has delivered a few samples.  Every rank
            # must run the same number of steps (fn contains collectives): rank 0 decides, the others follow.
            t_end, it2 = time.time() + 4.0, warmup + steps
            while True:
                more = smp is not None and len(smp.lines) < 5 and time.time() < t_end
                if dist is not None:
                    flag = torch.tensor([1 if more else 0], device=dev)
                    dist.broadcast(flag, 0)
                    more = bool(int(flag.item()))
                if not more:
                    break
                for _ in range(20):
                    fn(it2)
                    it2 += 1
                torch.cuda.synchronize()
            if smp:
                clocks = smp.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item()), clocks

    steps, warmup = args.steps, max(args.warmup, 3)
    # untimed pre-warm, both arms alike: every one of the 8 views three times, so that the caching
    # allocator has seen every scratch size and the clocks have ramped before the W warm-up steps
    for it in range(max(3 * N_VIEWS // vpr, 3)):
        step_resident(it)
    torch.cuda.synchronize()
    reset_state()
    sampler = ClockSampler(local) if rank == 0 else None
    if os.environ.get('FGS_BENCH_DIAG'):  # diagnosis only: repeat the timed region, report each on stderr
        for r in range(4):
            ms_d, _ = timed(step_resident, steps, warmup)
            print(f'[diag] loop {r}: {ms_d / steps:.4f} ms/step', file=sys.stderr)
        for h2d, d2h in ((True, True), (False, True), (True, False), (False, False)):
            diag['h2d'], diag['d2h'] = h2d, d2h
            ms_d, _ = timed(step_e2e, steps, warmup)
            print(f'[diag] e2e h2d={h2d} d2h={d2h}: {ms_d / steps:.4f} ms/step', file=sys.stderr)
        diag['h2d'] = diag['d2h'] = True
    ms_total, clocks = timed(step_resident, steps, warmup, sample_clocks=True)
    ms_e2e = float('nan') if args.quick else timed(step_e2e, steps, max(warmup, 10))[0]
    extra = {}
    if use_fused and not args.quick:  # also report the strict drop-in path: diff_rasterize backward + separate FusedAdam.step
        ms_u, _ = timed(lambda it: step_resident(it, fused=False), steps, 3)
        ms_ue, _ = timed(lambda it: step_e2e(it, fused=False), steps, 3)
        extra['dropin_unfused'] = {'value': round(batch * 1000.0 * steps / ms_u, 3), 'ms_per_step': round(ms_u / steps, 4),
                                   'e2e': round(batch * 1000.0 * steps / ms_ue, 3), 'unit': 'iters/s',
                                   'note': 'diff_rasterize + FusedAdam.step, the API the reference Trainer.py calls '
                                           '(loss as in `loss_path`)'}
    # second BASELINE.json metric: render ms (inference path, rasterize(..., to_chw=True)), inputs resident
    reset_state()
    with torch.no_grad():
        det = [p.detach() for p in params]
        for it in range(0 if args.quick else 5):
            be.rasterize(*det, be.RasterizerSettings(*st_dev[it % N_VIEWS]), to_chw=True)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for it in range(1 if args.quick else 50):
            be.rasterize(*det, be.RasterizerSettings(*st_dev[it % N_VIEWS]), to_chw=True)
        r1.record()
        torch.cuda.synchronize()
        render_ms = r0.elapsed_time(r1) / (1 if args.quick else 50)
    ms_step = ms_total / steps
    value = batch * 1000.0 / ms_step
    e2e_value = batch * 1000.0 / (ms_e2e / steps)

    if use_fused:
        opt_path = 'fused backward+Adam (diff_rasterize_fused, FasterGSFused semantics)' + (
            '; deferred SH schedule: sh_rest Adam on a side stream under the next forward\'s geometry pass + sorts' if defer_sh else '')
    elif dp_mode.startswith('peer-rows'):
        opt_path = ('backward into a symmetric-memory gradient buffer (no sh_rest gradient), then over NVLink peer memory: compact '
                    'gradient sum (14 of 59 floats per Gaussian and peer; sh_rest gradients rebuilt by the owner) + Adam on the '
                    'rank\'s Gaussians + broadcast' + (' (NVLS multimem.ld_reduce / multimem.st)' if opt_obj.multicast else ' (peer loads / stores)')
                    + ('; two kernels: the five small tensors on the main stream, ' if opt_obj.split else '; one kernel, ')
                    + ('sh_rest on a side stream under the next forward\'s geometry pass + sorts' if opt_obj.overlap_sh_broadcast
                       else 'sh_rest on the same stream'))
    elif dp_mode in ('peer', 'peer-mc'):
        opt_path = ('backward into a symmetric-memory gradient buffer, then ONE kernel: gradient sum over the ranks + Adam on the '
                    'rank\'s 1/W slice + parameter broadcast (' + ('NVLS multimem.ld_reduce / multimem.st' if opt_obj.multicast else 'NVLink peer loads / stores') + ')')
    elif dp_mode == 'sharded':
        opt_path = 'backward, NCCL reduce-scatter, Adam on the rank\'s 1/W shard, NCCL all-gather'
    elif dp_mode in ('allreduce', 'flat', 'pipelined'):
        opt_path = f'backward, NCCL all-reduce ({dp_mode}), dense FusedAdam.step on every rank'
    else:
        opt_path = 'backward' + (f' x{vpr} (gradient accumulation)' if vpr > 1 else '') + ', then FusedAdam.step'
    loss_desc = ('0.8 L1 + 0.2 DSSIM' if args.yaml_loss else 'L1') + (': fused loss+gradient kernel (csrc/loss.cu, fgs_b200.loss)' if fused_loss else ': torch ops')
    out = {
        'metric': cfg['metric'], 'value': round(value, 3), 'unit': 'iters/s', 'n_gpus': n_ranks, 'steps': steps,
        'warmup': warmup, 'ms_per_step': round(ms_step, 4), 'higher_is_better': True,
        'scaling': 'strong' if args.batch else 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['workload'] + f'; fwd + loss + bwd + Adam per view, {batch} view(s) per optimizer step',
                   'n_gaussians': n_gauss, 'width': width, 'height': height, 'views_per_step': batch,
                   'parallelism': f'view-dp{world}' if world > 1 else 'single',
                   'optimizer_path': opt_path,
                   'adam_schedule': 'overlapped (side-stream Adam of invisible rows)' if (args.overlap_adam and mine) else 'default',
                   'loss_path': loss_desc,
                   'densification_info': 'accumulated every step' if args.yaml_loss else 'not collected',
                   'l2_policy': f'inputs_larger_than_L2 (params+moments+grads = {944 * n_gauss // 1_000_000} MB per step vs 126 MB L2)'},
        'e2e': {'value': None if args.quick else round(e2e_value, 3), 'unit': 'iters/s',
                'h2d_bytes_per_step': vpr * (3 * width * height + 64 + 12), 'd2h_bytes_per_step': 4,
                'note': 'per view: uint8 target image + camera (w2c, position) from pinned host memory; per step: loss scalar back'},
        'clocks': clocks,
        'render_ms_1080p' if args.config != 'c3' else 'render_ms': round(render_ms, 4),
    }
    out.update(extra)
    if impl == 'reference':
        out['impl'] = 'reference'
        out['cpu_baseline'] = None
        out['config']['reference'] = 'unmodified FasterGSCudaBackend (baseline/_ref) CUDA path on the same GPU'
    else:
        dp_probe['on'] = True
        if hasattr(opt_obj, 'timing'):
            opt_obj.timing, opt_obj._timings = True, []
        out.update(profile_mine(be, cfg, params, opt, st_dev, gt_dev, gt8_dev, dev, ms_step, use_fused, fused_loss, step_resident,
                                opt_obj is not None or dist is not None, vpr))
        dp_probe['on'] = False
        torch.cuda.synchronize()
        if dp_events and (dist is not None or vpr > 1):
            # GPU time from the end of the last backward to the end of the optimizer step (waiting for the slowest rank at
            # the barrier included): gradient exchange + Adam + parameter broadcast
            out['step_end_ms'] = round(sum(a.elapsed_time(b) for a, b in dp_events) / len(dp_events), 4)
        if dist is not None:
            # per rank: the rasterizer's own kernels, the end of the step and its phases (who waits for whom)
            mine_row = {'rank': rank, 'kernels_ms_sum': out.get('kernels_ms_sum'), 'step_end_ms': out.get('step_end_ms')}
            if hasattr(opt_obj, 'timing'):
                mine_row['phases_ms'] = opt_obj.timing_summary()
                opt_obj.timing = False
            rows_all = [None] * world
            dist.all_gather_object(rows_all, mine_row)
            out['per_rank'] = rows_all
        if rank == 0 and n_ranks == 1 and not args.no_cpu_baseline and not args.quick:
            out['cpu_baseline'] = cpu_baseline(with_c2_estimate=args.config == 'c2')
    if registered is not None:
        registered.close()
    if dist is not None:
        dist.destroy_process_group()
    return out if rank == 0 else None


def profile_mine(be, cfg, params, opt, st_dev, gt_dev, gt8_dev, dev, ms_step, use_fused, fused_loss, step_fn, distributed, vpr):
    """Per-kernel CUDA-event timing of a few extra (untimed) iterations -> roofline of the dominant
    kernel, launch count, scene statistics."""
    from fgs_b200 import _capi
    n_gauss, width, height = cfg['n'], cfg['w'], cfg['h']
    t_tiles = math.ceil(width / 16) * math.ceil(height / 12)
    reps = 5
    _capi.profile_enable(True)
    per = {}
    launches = 0
    for it in range(reps):
        step_fn(it)
        recs = _capi.profile_read()
        # kernels per profiled group: a sort = one onesweep kernel per 8-bit digit place (the first-digit histograms
        # are accumulated by the kernels that produce the keys); create_instances = main + large-footprint kernel
        per_group = {'depth sort': 4, 'tile sort': math.ceil(math.ceil(math.log2(max(t_tiles, 2))) / 8), 'create_instances': 2}
        launches = sum(per_group.get(name, 1) for name, _ in recs)
        sums = {}
        for name, ms in recs:
            sums[name] = sums.get(name, 0.0) + ms
        for name, ms in sums.items():
            per.setdefault(name, []).append(ms)
    _capi.profile_enable(False)
    kern = {k: sum(v) / reps / vpr for k, v in per.items()}  # ms per view per kernel group
    # scene statistics of view 0 (V, I) from the forward state
    from fgs_b200 import _C
    with torch.no_grad():
        r = _C.forward(*[p.detach() for p in params], *st_dev[0])
    I = int(r[5])
    V = None
    try:
        V = int(torch.frombuffer(bytearray(r[1][:8].cpu().numpy().tobytes()), dtype=torch.int32)[0])
    except Exception:
        pass
    P, T = width * height, t_tiles
    N = n_gauss
    alg = algorithmic_bytes(N, V or 0, I, P, T)
    dom = max((k for k in kern if 'side stream' not in k), key=kern.get)
    peak, peak_src = peak_hbm_gbs()
    dom_bytes = alg.get(dom + '_fused' if (use_fused and dom == 'preprocess_backward') else dom, alg.get(dom, 0))
    traffic = None
    try:  # DRAM bytes per launch of that kernel from the committed ncu capture (profiles/)
        tj = ROOT / 'profiles' / ('r2_traffic.json' if (ROOT / 'profiles' / 'r2_traffic.json').exists() else 'r1_traffic.json')
        traffic = json.loads(tj.read_text())['bytes_per_launch'].get(dom)
    except Exception:
        pass
    dom_ms = kern[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    iter_bytes = alg['iteration_fused'] if use_fused else alg['iteration']
    iter_achieved = iter_bytes * vpr / (ms_step * 1e-3) / 1e9
    kernel_ms_sum = sum(v for k, v in kern.items() if 'side stream' not in k) * vpr  # main-stream kernels only
    out = {
        'gpu_launches': launches + (1 if fused_loss else 0) * vpr,
        'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': round(achieved, 2), 'peak': peak, 'unit': 'GB/s',
                     'frac': round(achieved / peak, 4), 'traffic': traffic, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': dom_bytes, 'ms_per_launch': round(dom_ms, 4),
                     'share_of_step': round(dom_ms * vpr / ms_step, 3),
                     'note': 'blend kernels are FP32-issue bound, not HBM bound; see DESIGN.md (issue roofline in profiles/)'},
        'roofline_iteration': {'algorithmic_bytes': iter_bytes * vpr, 'achieved': round(iter_achieved, 2),
                               'peak': peak, 'unit': 'GB/s', 'frac': round(iter_achieved / peak, 4)},
        'kernels_ms': {k: round(v, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1])},
        'kernels_ms_sum': round(kernel_ms_sum, 4),
        'scene': {'N': N, 'V': V, 'I': I, 'P': P, 'T': T},
    }
    if distributed:
        # what the step spends outside the rasterizer's own kernels: loss, gradient exchange + Adam, host gaps
        out['exposed_outside_rasterizer_ms'] = round(ms_step - kernel_ms_sum, 4)
    return out


def cpu_baseline(with_c2_estimate=True):
    """The oracle port (oracle/torch_oracle.py) timed on the host cores.  SURVEY.md §8(d): the reference has no CPU
    path (Renderer.py:58-59), so the reported CPU number is the pure-PyTorch oracle on C1 (1 k Gaussians, 128x128:
    forward + L1 + autograd backward + Adam), NOT extrapolated.  `c2_estimate` is a separate, clearly labelled extra:
    a bounded shape-preserving sample of the benchmark workload scaled to the full size."""
    from oracle import torch_oracle as O
    from fgs_b200 import synthetic as syn
    host_cores = os.cpu_count() or 1
    # the oracle is many small vectorised torch ops: beyond ~16 threads the fork/join overhead dominates
    # (measured: 128 threads are 100x slower than 8 on this workload), so at most 16 host threads are used
    threads = min(host_cores, 16)
    torch.set_num_threads(threads)
    t_start = time.perf_counter()
    sc = syn.make_scene(1000, 0)
    cam = syn.make_camera(0, 128, 128)
    s = O.Settings(cam.w2c, cam.position, torch.zeros(3), SH_BASES, 128, 128, cam.focal_x, cam.focal_y,
                   cam.center_x, cam.center_y, cam.near_plane, cam.far_plane, False)
    gt = syn.make_target(128, 128)
    leaves = [p.clone().requires_grad_(True) for p in sc]
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in leaves]

    def c1_iteration(step):
        out = O.render(*leaves, s)
        (out.image - gt).abs().mean().backward()
        with torch.no_grad():
            for p, (m, v), lr in zip(leaves, state, LRS):
                O.adam_step(p, p.grad, m, v, step, lr)
                p.grad = None

    c1_iteration(1)
    ts = []
    step = 2
    while len(ts) < 50 and (time.perf_counter() - t_start < 12.0 or len(ts) < 5):
        t0 = time.perf_counter()
        c1_iteration(step)
        ts.append(time.perf_counter() - t0)
        step += 1
    med = sorted(ts)[len(ts) // 2]
    out = {'value': round(1.0 / med, 4), 'unit': 'iters/s on C1 (1k Gaussians, 128x128)', 'cores': host_cores, 'threads': threads,
           'kind': 'port',
           'sample': f'oracle/torch_oracle.py fp32 on configs[0] (1 k Gaussians, 128x128 single view): forward + L1 + autograd '
                     f'backward + Adam, median of {len(ts)} iterations = {med * 1e3:.1f} ms, not extrapolated; '
                     f'{host_cores} host cores on this box, {threads} torch threads used'}
    if with_c2_estimate:
        out['c2_estimate'] = cpu_c2_estimate(threads)
    return out


def cpu_c2_estimate(threads):
    """EXTRA (an estimate, not the reported baseline): the benchmark scene scales Gaussian size with sqrt(1e6 / N)
    (fgs_b200/synthetic.py), so N/16 Gaussians rendered at 480x270 have the same footprint in pixels and the same depth
    complexity per pixel as the full 1M / 1080p workload.  Measured: preprocess forward+backward on the 62.5k Gaussians,
    blend forward+backward on a 96x72 window, Adam on the full 59 M parameters; scaled by 16, 16 x pixel ratio, 1."""
    from oracle import torch_oracle as O
    from fgs_b200 import synthetic as syn
    cfg = CONFIGS['c2']
    shrink = 16
    n, w, h = cfg['n'] // shrink, cfg['w'] // 4, cfg['h'] // 4
    sc = syn.make_scene(n, 0)
    cam = syn.make_camera(0, w, h)
    s = O.Settings(cam.w2c, cam.position, torch.zeros(3), SH_BASES, w, h, cam.focal_x, cam.focal_y,
                   cam.center_x, cam.center_y, cam.near_plane, cam.far_plane, False)
    win = (w // 2 - 48, h // 2 - 36, w // 2 + 48, h // 2 + 36)
    t_start = time.perf_counter()
    reps = 3

    def median_time(fn):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[len(ts) // 2]

    def run_pre():
        leaves = [p.clone().requires_grad_(True) for p in sc]
        pr = O.preprocess(*leaves, s)
        (pr.mean2d.sum() + pr.conic.sum() + pr.opacity.sum() + pr.color.sum()).backward()
    t_pre = median_time(run_pre)
    gt = syn.make_target(w, h)[:, win[1]:win[3], win[0]:win[2]]

    def run_all():
        leaves = [p.clone().requires_grad_(True) for p in sc]
        out = O.render(*leaves, s, window=win)
        (out.image - gt).abs().mean().backward()
    t_all = median_time(run_all)
    t_blend = max(t_all - t_pre, 0.0)
    full = [torch.zeros(cfg['n'], k) for k in (3, 3, 4, 1, 3, 45)]
    state = [(torch.ones_like(p), torch.zeros_like(p), torch.zeros_like(p)) for p in full]

    def run_adam():
        with torch.no_grad():
            for p, (g, m, v), lr in zip(full, state, LRS):
                O.adam_step(p, g, m, v, 1, lr)
    t_adam = median_time(run_adam)
    ratio = (w * h) / ((win[2] - win[0]) * (win[3] - win[1]))
    est = shrink * t_pre + shrink * ratio * t_blend + t_adam
    return {'iters_per_s': round(1.0 / est, 6), 'kind': 'ESTIMATE (scaled bounded sample; not the reported baseline)',
            'sample': f'1/{shrink} shape-preserving sample: {n} Gaussians at {w}x{h}; preprocess fwd+bwd {t_pre:.1f} s (x{shrink}), '
                      f'blend fwd+bwd on a 96x72 window {t_blend:.1f} s (x{shrink * ratio:.0f}), Adam on 59M params {t_adam:.1f} s '
                      f'(medians of {reps}); {time.perf_counter() - t_start:.0f} s of CPU work, estimated {est:.0f} s per iteration'}


def run_cpu_port(args, reason):
    cb = cpu_baseline(with_c2_estimate=False)
    return {'metric': CONFIGS[args.config]['metric'], 'value': cb['value'], 'unit': 'iters/s', 'n_gpus': 0, 'steps': 1, 'warmup': 0,
            'ms_per_step': round(1000.0 / cb['value'], 1), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': 'configs[0] (C1: the CPU oracle port; the reference has no CPU path)', 'note': reason},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': 'iters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}


def run_fourd(args):
    """configs[4]: FasterGS4D shape.  One step = time slice + forward + L1 + backward + Adam on the seven 4D tensors at a
    random time stamp; render ms = inference over the 150 time steps.  THIS repo only (the reference's FasterGS4D branch
    is not in the tree)."""
    import fgs_b200
    from fgs_b200 import fourd as F4, loss as loss_mod, synthetic as syn
    from tests.test_widen_fourd import make_scene_4d
    cfg = CONFIGS['c5']
    dev = torch.device('cuda', 0)
    n, w, h, n_t = cfg['n'], cfg['w'], cfg['h'], 150
    params = [p.to(dev).requires_grad_(True) for p in make_scene_4d(n, 0, w)]
    lrs4 = [1.6e-4, 5e-3, 1e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]
    opt = fgs_b200.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs4)], lr=0.0, eps=1e-15)
    cams = [syn.make_camera(k, w, h) for k in range(N_VIEWS)]
    bg = torch.zeros(3, device=dev)
    rs = [fgs_b200.RasterizerSettings(*settings_tuple(c, dev, bg)) for c in cams]
    targets8 = [(syn.make_target(w, h, seed=k) * 255.0).round().clamp(0, 255).to(torch.uint8) for k in range(N_VIEWS)]
    gt8 = [t.to(dev) for t in targets8]
    gt_pin = [t.pin_memory() for t in targets8]
    times = [i / (n_t - 1) for i in range(n_t)]
    empty = torch.empty(0)

    def step(i, gt):
        img = F4.diff_rasterize_4d(*params, times[(37 * i) % n_t], empty, rs[i % N_VIEWS])
        terms, g = loss_mod.l1_dssim_loss_and_grad(img.detach(), gt, 1.0, 0.0)
        img.backward(g)
        opt.step()
        opt.zero_grad()
        return terms[0]

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    steps, warmup = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(0)
    sampler.start()
    ms_step = timed(lambda i: step(i, gt8[i % N_VIEWS]), steps, warmup)
    stage = [torch.empty_like(gt8[0]) for _ in range(2)]
    loss_pin = torch.zeros(1).pin_memory()

    def step_e2e(i):
        b = stage[i % 2]
        b.copy_(gt_pin[i % N_VIEWS], non_blocking=True)
        loss = step(i, b)
        loss_pin.copy_(loss.reshape(1), non_blocking=True)
    ms_e2e = timed(step_e2e, steps, warmup)
    with torch.no_grad():
        det = [p.detach() for p in params]
        ms_render = timed(lambda i: F4.rasterize_4d(*det, times[i % n_t], rs[i % N_VIEWS], to_chw=True), n_t, 10)
    t_end = time.time() + 3.0
    while len(sampler.lines) < 5 and time.time() < t_end:
        for i in range(20):
            step(i, gt8[i % N_VIEWS])
        torch.cuda.synchronize()
    clocks = sampler.stop()
    peak, peak_src = peak_hbm_gbs()
    # algorithmic bytes of the 4D iteration: the 3D iteration on the sliced Gaussians (unfused: SURVEY §8d) with the slice
    # forward (25 floats in, 10 out per Gaussian) and backward (25 + 10 in, 25 out) and Adam on 4+4+4+4+1+3+45 floats
    adam_floats = 4 + 4 + 4 + 4 + 1 + 3 + 45
    return {'metric': cfg['metric'], 'value': round(1000.0 / ms_step, 3), 'unit': 'iters/s', 'n_gpus': 1, 'steps': steps, 'warmup': warmup,
            'ms_per_step': round(ms_step, 4), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': cfg['workload'] + '; slice + fwd + fused L1 + bwd + FusedAdam.step on the seven 4D tensors, random time stamp',
                       'n_gaussians': n, 'width': w, 'height': h, 'time_steps': n_t,
                       'reference': 'none: the FasterGS4D branch is not in /root/reference (parity unpinned, oracle/fourd_oracle.py)',
                       'l2_policy': f'inputs_larger_than_L2 (params+moments+grads = {adam_floats * 16 * n // 1_000_000} MB per step vs 126 MB L2)'},
            'e2e': {'value': round(1000.0 / ms_e2e, 3), 'unit': 'iters/s', 'h2d_bytes_per_step': 3 * w * h, 'd2h_bytes_per_step': 4},
            'render_ms_per_timestep': round(ms_render, 4), 'clocks': clocks, 'gpu_launches': 22,
            'roofline': {'bound': 'hbm', 'kernel': 'whole iteration (no per-kernel split for the 4D path)',
                         'achieved': None, 'peak': peak, 'unit': 'GB/s', 'frac': None, 'traffic': None, 'peak_source': peak_src},
            'cpu_baseline': None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='mine', choices=['mine', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS), help='c2 = BASELINE.json configs[1] (headline), c3 = configs[2], c5 = configs[4]')
    ap.add_argument('--batch', type=int, default=0,
                    help='views per optimizer step over the whole job (configs[3]: --batch 8 --gpus 1/2/4/8, strong scaling); '
                         'default: one view per rank (weak scaling)')
    ap.add_argument('--dp-mode', default='auto', choices=DP_MODES,
                    help='N > 1: how gradients are summed and Adam applied.  peer-mc / peer = one kernel over NVLS multicast / '
                         'NVLink peer memory (fgs_b200.parallel.PeerAdam); sharded = NCCL reduce-scatter + Adam on 1/W + '
                         'all-gather; allreduce / flat / pipelined = NCCL all-reduce (six / one / six pipelined with Adam) + dense Adam')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--quick', action='store_true', help='experiments only: resident timing + kernel profile, no e2e / drop-in / render / CPU legs')
    ap.add_argument('--unfused', action='store_true', help='headline on the drop-in path (backward, then FusedAdam.step)')
    ap.add_argument('--torch-loss', action='store_true', help='loss with torch ops ((img - gt).abs().mean()) instead of the fused loss kernel')
    ap.add_argument('--yaml-loss', action='store_true',
                    help='the loss fastergs_garden.yaml trains with (0.8 L1 + 0.2 DSSIM, Loss.py:15-16) and a live densification_info')
    ap.add_argument('--defer-sh', action='store_true',
                    help='fused path with the deferred SH schedule (sh_rest Adam on a side stream under the next forward\'s geometry '
                         'pass; measured +0.5 % at C2, so not the default)')
    ap.add_argument('--no-defer-sh', action='store_true', help='(default; kept for compatibility)')
    ap.add_argument('--small-mc', default='auto', choices=['auto', 'on', 'off'],
                    help='N > 1, peer-rows: the small-tensor kernel through NVLS multimem (auto: at 8 ranks)')
    ap.add_argument('--side-ctas', type=int, default=0, help='experiments: CTAs per SM of the side-stream sh_rest kernel at N > 1 (default 2)')
    ap.add_argument('--defer-ctas', type=int, default=0, help='experiments: CTAs per SM of the deferred sh_rest update kernel (default 2)')
    ap.add_argument('--overlap-adam', action='store_true',
                    help='MEASURED SLOWER (profiles/r2_results.md), kept for reference: zero-gradient Adam of invisible rows on a '
                         'side stream under the blend backward')
    # round-1 spellings
    ap.add_argument('--fused-loss', action='store_true', help='(default now; kept for compatibility)')
    ap.add_argument('--flat-allreduce', action='store_true', help='= --dp-mode flat')
    ap.add_argument('--peer-adam', action='store_true', help='= --dp-mode peer')
    args = ap.parse_args()
    if args.flat_allreduce:
        args.dp_mode = 'flat'
    if args.peer_adam:
        args.dp_mode = 'peer'
    if not torch.cuda.is_available():
        if args.impl == 'reference':
            print(json.dumps(run_cpu_port(args, 'no CUDA device')))
            return
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    if args.config == 'c5':
        if int(os.environ.get('RANK', 0)) != 0:
            return
        if args.impl == 'reference':
            print(json.dumps({'impl': 'reference', 'unavailable': 'configs[4]: the FasterGS4D branch is not in /root/reference'}))
            return
        print(json.dumps(run_fourd(args)))
        return
    out = run_gpu(args, args.impl)
    if out is not None:
        print(json.dumps(out))


if __name__ == '__main__':
    main()
