"""View-parallel data parallelism for multi-view batches (new functionality -- the
reference trains one view per iteration on one GPU, Trainer.py:180, Renderer.py:58-61).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The path shards over VIEWS only:
every rank holds a full replica of the Gaussians, renders / back-propagates its own views of the
batch, then the six parameter gradients (236 bytes per Gaussian) -- and the densification statistics
while they are being collected -- are summed across ranks and every rank applies the identical Adam
step.  The 1-GPU semantic reference for a batch of B views is gradient accumulation over the B views
followed by one FusedAdam.step, so gradients are SUMMED, not averaged.  Nothing else is exchanged:
sorting and blending of a single view are never split across GPUs.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch


def shard_views(n_views: int, rank: int, world_size: int) -> list[int]:
    """Views of the batch rendered by `rank`: {rank, rank + W, rank + 2W, ...}."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of size {world_size}')
    return list(range(rank, n_views, world_size))


def allreduce_gradients(params: Sequence[torch.Tensor], densification_info: torch.Tensor | None = None,
                        group=None) -> None:
    """Sum .grad of every parameter (and densification_info, if given) over the process group.
    All collectives are launched asynchronously and waited for together, so NCCL can pipeline them."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    works = []
    for p in params:
        if p.grad is None:
            raise RuntimeError('allreduce_gradients: parameter without gradient (every rank must render a view)')
        works.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
    if densification_info is not None and densification_info.numel() > 0:
        works.append(dist.all_reduce(densification_info, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()


def allreduce_and_step(params: Sequence[torch.Tensor], optimizer, densification_info: torch.Tensor | None = None,
                       group=None, step_group=None) -> None:
    """`allreduce_gradients` followed by `optimizer.step()`, pipelined: the all-reduces are launched largest tensor
    first (the same order on every rank) and each tensor's Adam update is queued as soon as ITS all-reduce has
    completed, so the optimizer kernels of the big tensors run while NCCL is still reducing the small ones.
    Same result as the two separate calls (the update of a tensor only depends on that tensor's summed gradient).

    `optimizer` is a FusedAdam (one tensor per param group); `step_group(param_group)` overrides how one group is
    stepped (used by the CPU tests)."""
    import torch.distributed as dist
    step_one = step_group if step_group is not None else optimizer.step_group
    group_of = {id(g['params'][0]): g for g in optimizer.param_groups}
    for p in params:
        if id(p) not in group_of:
            raise KeyError('allreduce_and_step: parameter is not managed by the optimizer')
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    order = sorted(range(len(params)), key=lambda i: -params[i].numel())  # stable: ties keep the caller's order
    works = []
    if distributed:
        for i in order:
            if params[i].grad is None:
                raise RuntimeError('allreduce_and_step: parameter without gradient (every rank must render a view)')
            works.append(dist.all_reduce(params[i].grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
        if densification_info is not None and densification_info.numel() > 0:
            works.append(dist.all_reduce(densification_info, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for k, i in enumerate(order):
        if distributed:
            works[k].wait()  # stream-ordered: the compute stream waits for this collective only
        step_one(group_of[id(params[i])])
    for w in works[len(order):]:
        w.wait()
    managed = {id(p) for p in params}
    for g in optimizer.param_groups:  # groups the caller did not hand in: plain local step
        if id(g['params'][0]) not in managed:
            step_one(g)


def accumulate_views(render_and_backward, views: Iterable[int]) -> int:
    """Gradient accumulation over this rank's views: calls `render_and_backward(view)` for each, which
    must run forward + loss.backward() so that .grad accumulates.  Returns the number of views."""
    n = 0
    for v in views:
        render_and_backward(v)
        n += 1
    return n


# ---- reduce + Adam + broadcast over NVLink peer memory (measured on 2 / 4 / 8 B200s: profiles/r2_results.md) ------------
def peer_layout(shapes: Sequence[Sequence[int]]) -> tuple[list[int], list[int], int]:
    """Flat layout shared by every rank: (offsets, counts, total) in floats; every tensor region starts at a
    multiple of 4 floats and is padded to one (include/fgs_b200.h, fgs_peer_adam_step)."""
    offsets, counts, off = [], [], 0
    for shape in shapes:
        n = 1
        for d in shape:
            n *= int(d)
        offsets.append(off)
        counts.append(n)
        off += (n + 3) // 4 * 4
    return offsets, counts, off


class FlatGradients:
    """One contiguous buffer behind the six gradient tensors, so that the view-parallel reduction is ONE all-reduce of
    236 N bytes instead of six collectives (five of them latency-bound).  While the object is alive the rasterizer's
    backward writes its gradients straight into views of the buffer (`_C.set_gradient_outputs`; autograd adopts them as
    `.grad` without a copy), so nothing is gathered or scattered.  One backward per rank and step: a second backward before
    `zero_grad()` would overwrite the buffer the first one's `.grad` lives in (accumulate views with separate tensors).

        flat = FlatGradients(params)                       # every rank
        loss.backward(); flat.allreduce(params, dens); optimizer.step(); optimizer.zero_grad()

    EXPERIMENTAL: covered by the gloo tests, not yet measured over NCCL (bench.py --flat-allreduce)."""

    def __init__(self, params: Sequence[torch.Tensor], register: bool = True) -> None:
        self.shapes = [tuple(p.shape) for p in params]
        self.offsets, self.counts, total = peer_layout(self.shapes)
        self.buffer = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        self.views = [self.buffer[o:o + c].view(shape) for o, c, shape in zip(self.offsets, self.counts, self.shapes)]
        self._registered = False
        if register:
            from . import _C
            _C.set_gradient_outputs(self.views)
            self._registered = True

    def allreduce(self, params: Sequence[torch.Tensor], densification_info: torch.Tensor | None = None, group=None) -> None:
        """Sum the gradients (and densification_info) over the group; afterwards every p.grad is a view of the buffer."""
        import torch.distributed as dist
        for p, v in zip(params, self.views):
            if p.grad is None:
                raise RuntimeError('FlatGradients.allreduce: parameter without gradient (every rank must render a view)')
            if p.grad.data_ptr() != v.data_ptr():  # produced elsewhere (another loss term, the CPU tests): bring it in
                v.copy_(p.grad)
                p.grad = v
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        works = [dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group, async_op=True)]
        if densification_info is not None and densification_info.numel() > 0:
            works.append(dist.all_reduce(densification_info, op=dist.ReduceOp.SUM, group=group, async_op=True))
        for w in works:
            w.wait()

    def close(self) -> None:
        if self._registered:
            from . import _C
            _C.set_gradient_outputs(None)
            self._registered = False


class PeerAdam:
    """Dense Adam over a view-parallel group without an all-reduce: gradients and parameters live in symmetric memory
    (`torch.distributed._symmetric_memory`), one kernel per step (csrc/peer_adam.cu) sums this rank's slice of every
    gradient tensor over all peers, applies Adam with the rank's shard of the moments and stores the new parameters
    to every peer.  Same result as `allreduce_gradients` + `FusedAdam.step` (gradients SUMMED over ranks), replicas
    bit-identical by construction.

        opt = PeerAdam(initial_parameter_tensors, lrs)       # on every rank, identical values
        img = diff_rasterize(*opt.params, dens, settings); loss(img).backward(); opt.step(); opt.zero_grad()

    `opt.params` are leaf tensors viewing the symmetric parameter buffer; while the object is alive, the
    rasterizer's backward writes its gradients straight into the symmetric gradient buffer."""

    def __init__(self, tensors: Sequence[torch.Tensor], lrs: Sequence[float], betas=(0.9, 0.999), eps: float = 1e-15,
                 group=None, multicast: bool | str = 'auto', compact: bool = False, active_sh_bases: int = 16,
                 overlap_sh_broadcast: bool = True, split: bool = True, small_multicast: bool | str = 'auto') -> None:
        """`multicast`: True = reduce and broadcast inside the NVSwitch through the buffers' multicast addresses
        (multimem.ld_reduce / multimem.st, fgs_peer_adam_step_multicast), False = W peer loads / stores per element,
        'auto' = multicast when the symmetric-memory handles expose a multicast address.

        `compact` (one view per rank and step): ranks own Gaussians (rows) and the peers' sh_rest gradients are never
        exchanged -- the owner rebuilds them from the peers' sh0 gradients and camera positions
        (fgs_peer_adam_rows_step): 14 instead of 59 floats per Gaussian and peer.  The backward then does not even write
        the sh_rest gradient (`p.grad` of sh_rest stays None).  With `overlap_sh_broadcast` the new sh_rest parameters
        (three quarters of the parameter bytes) travel to the peers on a side stream, by copy engines, while the next
        forward's geometry pass and sorts run; that forward must be `forward_settings()`-aware: pass
        `color_event=opt.sh_ready` and `defer_color=True` (PeerAdam installs that for `diff_rasterize` itself).

        `split` (default, with `compact`): two kernels instead of one.  The five small tensors (56 of the 236 bytes of a
        Gaussian) go through the float4-sliced kernel at full occupancy on the caller's stream -- the only part the next
        forward's geometry pass has to wait for; sh_rest (rebuild the gradient from the peers' sh0 gradients, Adam,
        broadcast to the peers) is a second kernel on the side stream.  `split=False`: one row-sliced kernel does
        everything and only the broadcast of the new sh_rest parameters is overlapped (measured slower: the row-sliced
        kernel is a chain of remote round trips at 12 warps per SM)."""
        import ctypes as C
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _C, _capi
        if len(tensors) != 6 or len(lrs) != 6:
            raise ValueError('PeerAdam: six tensors / learning rates (means, scales, rotations, opacities, sh0, sh_rest)')
        self._C, self._capi, self._ctypes = _C, _capi, C
        self._lib = _capi.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        dev = tensors[0].device
        self.shapes = [tuple(t.shape) for t in tensors]
        self.offsets, self.counts, total = peer_layout(self.shapes)
        self.lrs, self.betas, self.eps, self.steps = [float(x) for x in lrs], betas, float(eps), 0
        self._cam_off = total  # 4 floats behind the six gradient regions: this rank's camera position (compact exchange)
        self._pbuf = symm.empty(total, dtype=torch.float32, device=dev)
        self._gbuf = symm.empty(total + 4, dtype=torch.float32, device=dev)
        self._pbuf.zero_()
        self._gbuf.zero_()
        self._ph = symm.rendezvous(self._pbuf, self.group.group_name)
        self._gh = symm.rendezvous(self._gbuf, self.group.group_name)
        view = lambda buf, i: buf[self.offsets[i]:self.offsets[i] + self.counts[i]].view(self.shapes[i])
        with torch.no_grad():
            for i, t in enumerate(tensors):
                view(self._pbuf, i).copy_(t)
        self.params = [view(self._pbuf, i).requires_grad_(True) for i in range(6)]
        self.grads = [view(self._gbuf, i) for i in range(6)]
        self.compact, self.active_sh_bases, self.split = bool(compact), int(active_sh_bases), bool(split) and bool(compact)
        self.overlap_sh_broadcast = bool(overlap_sh_broadcast) and self.compact and self.world > 1
        self.n_rows = int(self.shapes[0][0])
        self.n_rest = int(self.shapes[5][1]) if len(self.shapes[5]) > 1 else 0
        if self.compact:
            row_floats = [3, 3, 4, 1, 3, 3 * self.n_rest]
            if [c for c in self.counts] != [self.n_rows * k for k in row_floats]:
                raise ValueError('PeerAdam(compact=True): tensors must be (N,3) (N,3) (N,4) (N,1) (N,1,3) (N,R,3)')
            self.rows_per_rank = int(self._lib.fgs_peer_adam_rows_per_rank(self.n_rows, self.world))
            shard = self.rows_per_rank * sum(row_floats)
        else:
            shard = int(self._lib.fgs_peer_adam_shard_floats((C.c_uint64 * 6)(*self.counts), self.world))
        self.exp_avg = torch.zeros(shard, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(shard, dtype=torch.float32, device=dev)
        self._gptrs = (C.c_void_p * self.world)(*[int(p) for p in self._gh.buffer_ptrs])
        self._pptrs = (C.c_void_p * self.world)(*[int(p) for p in self._ph.buffer_ptrs])
        g_mc, p_mc = int(getattr(self._gh, 'multicast_ptr', 0) or 0), int(getattr(self._ph, 'multicast_ptr', 0) or 0)
        if multicast == 'auto':
            multicast = g_mc != 0 and p_mc != 0
        if multicast and not (g_mc and p_mc):
            raise RuntimeError('PeerAdam(multicast=True): the symmetric-memory handles have no multicast address (no NVLS)')
        self.multicast, self._g_mc, self._p_mc = bool(multicast), g_mc, p_mc
        if self.compact:
            # NVLS for the compact step: plain sums via multimem.ld_reduce, broadcasts via multimem.st (K1), and the
            # overlapped sh_rest broadcast as one small multimem.st kernel instead of W - 1 copy-engine transfers
            self.multicast = bool(self.multicast and g_mc and p_mc)
            # the backward skips the sh_rest gradient altogether and leaves the view's camera position in the slot
            _C.set_gradient_outputs(self.grads[:5] + [None], camera_out=self._gbuf[self._cam_off:self._cam_off + 3])
            self._side = torch.cuda.Stream(device=dev)
            self._k1_done = torch.cuda.Event()
            self.sh_ready = torch.cuda.Event()  # all peers' sh_rest slices have landed in this replica
            self.sh_ready.record()
            lo = min(self.rank * self.rows_per_rank, self.n_rows)
            hi = min(lo + self.rows_per_rank, self.n_rows)
            k = 3 * self.n_rest
            self._rest_lo, self._rest_hi = self.offsets[5] + lo * k, self.offsets[5] + hi * k
            self._peer_pbufs = [self._ph.get_buffer(q, (total,), torch.float32) for q in range(self.world)]
            if self.split:
                self.multicast = False  # the sh_rest kernel: peer stores (ingress-bound either way, multimem only adds latency)
                # the small-tensor kernel is NVLink-bound at 8 ranks (98 MB each way): there the switch's reduction and
                # replication (56 MB each way) pays, below that multimem's latency does not
                if small_multicast == 'auto':
                    small_multicast = False  # the measured default; `True` is the experiment (bench.py --small-mc on)
                if small_multicast and not (g_mc and p_mc):
                    raise RuntimeError('PeerAdam(small_multicast=True): the symmetric-memory handles have no multicast address (no NVLS)')
                self.small_multicast = bool(small_multicast)
                self._means_old = torch.zeros(self.rows_per_rank * 3 + 4, dtype=torch.float32, device=dev)
            if self.overlap_sh_broadcast:  # every diff_rasterize forward from now on defers its colour kernel behind sh_ready
                from .torch_bindings import rasterization as _rast
                _rast.set_forward_kwargs(self.forward_kwargs)
        else:
            _C.set_gradient_outputs(self.grads)
        self._ph.barrier()  # every replica initialised before anyone steps

    def step(self) -> None:
        C = self._ctypes
        for i, (p, g) in enumerate(zip(self.params, self.grads)):
            if self.compact and i == 5:
                continue  # rebuilt by the owner from the sh0 gradients
            if p.grad is None:
                raise RuntimeError('PeerAdam.step: parameter without gradient (every rank must render a view)')
            if p.grad.data_ptr() != g.data_ptr():  # gradient produced elsewhere (e.g. a regulariser): bring it in
                g.copy_(p.grad)
        self.steps += 1
        stream = self._capi.current_stream()
        ev = self._mark('enter')
        self._gh.barrier()  # stream-ordered: every rank's gradients are complete
        self._mark('gradients complete on every rank', ev)
        if self.compact:
            self._step_compact(stream, ev)
            return
        tail = (self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), (C.c_uint64 * 6)(*self.offsets),
                (C.c_uint64 * 6)(*self.counts), (C.c_int32 * 6)(*[self.steps] * 6), (C.c_double * 6)(*self.lrs),
                float(self.betas[0]), float(self.betas[1]), self.eps, stream)
        if self.multicast:
            rc = self._lib.fgs_peer_adam_step_multicast(self._g_mc, self._p_mc, self._pbuf.data_ptr(), self.world, self.rank, *tail)
        else:
            rc = self._lib.fgs_peer_adam_step(self._gptrs, self._pptrs, self.world, self.rank, *tail)
        if rc != 0:
            raise RuntimeError(self._lib.fgs_peer_last_error().decode())
        self._mark('kernel', ev)
        self._ph.barrier()  # every rank's parameter stores have landed before the next forward reads them
        self._mark('parameters landed everywhere', ev)

    broadcast_ctas = 64  # grid of the overlapped sh_rest broadcast kernel (NVLS): small, it runs next to the forward

    # ---- optional timing of the phases of a step (bench.py): CUDA events on the streams the work is queued on
    timing = False
    _timings: list = []

    def _mark(self, name, ev=None):
        if not self.timing:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if ev is None:
            ev = []
            self._timings.append(ev)
        ev.append((name, e))
        return ev

    def timing_summary(self) -> dict:
        """Mean milliseconds per phase over the recorded steps: waiting for the slowest rank (first barrier), the
        kernel, the closing barrier; for the overlapped sh_rest broadcast also the side stream's copies + barrier."""
        torch.cuda.synchronize()
        out: dict = {}
        for ev in self._timings:
            for (_, a), (name, b) in zip(ev, ev[1:]):
                out.setdefault(name, []).append(a.elapsed_time(b))
        self._timings = []
        return {k: round(sum(v) / len(v), 4) for k, v in out.items()}

    side_ctas_per_sm = 2  # grid bound of the sh_rest kernel when it runs on the side stream, next to the forward

    def _step_split(self, stream: int, ev=None) -> None:
        C = self._ctypes
        common = ((C.c_int32 * 6)(*[self.steps] * 6), (C.c_double * 6)(*self.lrs), float(self.betas[0]), float(self.betas[1]), self.eps)
        rc = self._lib.fgs_peer_adam_small_step(self._gptrs, self._pptrs, self.world, self.rank, self.exp_avg.data_ptr(),
                                                self.exp_avg_sq.data_ptr(), (C.c_uint64 * 6)(*self.offsets), self.n_rows,
                                                self.n_rest, *common, self._means_old.data_ptr(),
                                                self._g_mc if self.small_multicast else None,
                                                self._p_mc if self.small_multicast else None, stream)
        if rc != 0:
            raise RuntimeError(self._lib.fgs_peer_last_error().decode())
        self._mark('kernel', ev)

        def rest(on_stream: int, ctas: int) -> None:
            rc2 = self._lib.fgs_peer_adam_rows_step(self._gptrs, self._pptrs, self.world, self.rank, self.exp_avg.data_ptr(),
                                                    self.exp_avg_sq.data_ptr(), (C.c_uint64 * 6)(*self.offsets), self.n_rows,
                                                    self.n_rest, self.active_sh_bases, self._cam_off, 1, *common, None, None,
                                                    self._means_old.data_ptr(), ctas, on_stream)
            if rc2 != 0:
                raise RuntimeError(self._lib.fgs_peer_last_error().decode())

        if self.overlap_sh_broadcast:
            self._k1_done.record()
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._k1_done)
                sev = self._mark('side: start')
                rest(self._capi.current_stream(), self.side_ctas_per_sm)
                self._mark('side: sh_rest kernel (rebuild gradient + Adam + broadcast)', sev)
                self._ph.barrier(channel=1)
                self._mark('side: barrier', sev)
                self.sh_ready.record()
        else:
            rest(stream, 0)
            self._mark('sh_rest kernel', ev)
        self._ph.barrier()  # the five small tensors of every rank's Gaussians have landed everywhere
        self._mark('parameters landed everywhere', ev)
        if not self.overlap_sh_broadcast:
            self.sh_ready.record()

    def _step_compact(self, stream: int, ev=None) -> None:
        if self.split:
            self._step_split(stream, ev)
            return
        C = self._ctypes
        overlap = self.overlap_sh_broadcast
        rc = self._lib.fgs_peer_adam_rows_step(self._gptrs, self._pptrs, self.world, self.rank, self.exp_avg.data_ptr(),
                                               self.exp_avg_sq.data_ptr(), (C.c_uint64 * 6)(*self.offsets), self.n_rows,
                                               self.n_rest, self.active_sh_bases, self._cam_off, 0 if overlap else 1,
                                               (C.c_int32 * 6)(*[self.steps] * 6), (C.c_double * 6)(*self.lrs),
                                               float(self.betas[0]), float(self.betas[1]), self.eps,
                                               self._g_mc if self.multicast else None, self._p_mc if self.multicast else None,
                                               None, 0, stream)
        if rc != 0:
            raise RuntimeError(self._lib.fgs_peer_last_error().decode())
        self._mark('kernel', ev)
        if overlap:
            # this rank's slice of the new sh_rest parameters goes to the peers on the side stream (copy engines over NVLink)
            # while the caller's stream continues with the next forward's geometry pass; `sh_ready` = all slices are here
            self._k1_done.record()
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._k1_done)
                sev = self._mark('side: start')
                if self._rest_hi > self._rest_lo and self.multicast:
                    rc = self._lib.fgs_peer_broadcast_multicast(self._pbuf.data_ptr(), self._p_mc, self._rest_lo,
                                                                self._rest_hi - self._rest_lo, self.broadcast_ctas,
                                                                self._capi.current_stream())
                    if rc != 0:
                        raise RuntimeError(self._lib.fgs_peer_last_error().decode())
                elif self._rest_hi > self._rest_lo:
                    src = self._pbuf[self._rest_lo:self._rest_hi]
                    for q in range(self.world):
                        if q != self.rank:
                            self._peer_pbufs[q][self._rest_lo:self._rest_hi].copy_(src, non_blocking=True)
                self._mark('side: sh_rest copies to the peers', sev)
                self._ph.barrier(channel=1)
                self._mark('side: barrier', sev)
                self.sh_ready.record()
        self._ph.barrier()  # the geometry parameters (and sh0) of every rank's slice have landed everywhere
        self._mark('parameters landed everywhere', ev)
        if not overlap:
            self.sh_ready.record()

    def forward_kwargs(self) -> dict:
        """Keyword arguments for `_C.forward_ex` / `diff_rasterize_dp` of the NEXT forward: with the overlapped sh_rest
        broadcast its colour kernel has to wait for `sh_ready`."""
        if self.compact and self.overlap_sh_broadcast:
            return {'defer_color': True, 'color_event': self.sh_ready}
        return {}

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    def reset(self) -> None:
        """Fresh optimizer state (moments and step count); the parameters are left alone."""
        if self.compact:
            torch.cuda.current_stream().wait_event(self.sh_ready)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.steps = 0

    def close(self) -> None:
        if self.compact:
            torch.cuda.current_stream().wait_event(self.sh_ready)
            if self.overlap_sh_broadcast:
                from .torch_bindings import rasterization as _rast
                _rast.set_forward_kwargs(None)
        self._C.set_gradient_outputs(None)


class ShardedAdam:
    """The same step with NCCL collectives instead of peer memory: reduce-scatter of the flat gradient buffer, dense
    Adam on this rank's 1/W of every tensor (moments sharded), all-gather of the flat parameter buffer.  Same wire bytes
    as one all-reduce, Adam's HBM traffic / W; the fallback where symmetric memory is not available, and the yardstick
    PeerAdam is measured against (bench.py --sharded-adam).

        opt = ShardedAdam(initial_parameter_tensors, lrs)
        img = diff_rasterize(*opt.params, dens, settings); loss(img).backward(); opt.step(); opt.zero_grad()
    """

    def __init__(self, tensors: Sequence[torch.Tensor], lrs: Sequence[float], betas=(0.9, 0.999), eps: float = 1e-15,
                 group=None, adam_fn=None, register: bool = True) -> None:
        """`adam_fn(grad, param, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps)` replaces the library's Adam kernel and
        `register=False` leaves the rasterizer's gradient outputs alone (both for the CPU tests over gloo)."""
        import torch.distributed as dist
        self.group = group
        self._C = None
        if adam_fn is None or register:
            from . import _C
            self._C = _C
        self._adam = adam_fn if adam_fn is not None else self._C.adam_step
        self._registered = False
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        dev = tensors[0].device
        self.shapes = [tuple(t.shape) for t in tensors]
        self.offsets, self.counts, total = peer_layout(self.shapes)
        self.chunk = (total + 4 * self.world - 1) // (4 * self.world) * 4  # floats per rank, multiple of 4
        padded = self.chunk * self.world
        self.lrs, self.betas, self.eps, self.steps = [float(x) for x in lrs], betas, float(eps), 0
        self._pbuf = torch.zeros(padded, dtype=torch.float32, device=dev)
        self._gbuf = torch.zeros(padded, dtype=torch.float32, device=dev)
        view = lambda buf, i: buf[self.offsets[i]:self.offsets[i] + self.counts[i]].view(self.shapes[i])
        with torch.no_grad():
            for i, t in enumerate(tensors):
                view(self._pbuf, i).copy_(t)
        self.params = [view(self._pbuf, i).requires_grad_(True) for i in range(6)]
        self.grads = [view(self._gbuf, i) for i in range(6)]
        lo, hi = self.rank * self.chunk, (self.rank + 1) * self.chunk
        self._gshard = torch.zeros(self.chunk, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.chunk, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.chunk, dtype=torch.float32, device=dev)
        # pieces of this rank's chunk, one per tensor it intersects: (shard offset, length, lr)
        self._pieces = []
        for o, c, lr in zip(self.offsets, self.counts, self.lrs):
            a, b = max(lo, o), min(hi, o + c)
            if b > a:
                self._pieces.append((a - lo, b - a, lr))
        if register:
            self._C.set_gradient_outputs(self.grads)
            self._registered = True

    def step(self) -> None:
        import torch.distributed as dist
        for p, g in zip(self.params, self.grads):
            if p.grad is None:
                raise RuntimeError('ShardedAdam.step: parameter without gradient (every rank must render a view)')
            if p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
        self.steps += 1
        dist.reduce_scatter_tensor(self._gshard, self._gbuf, op=dist.ReduceOp.SUM, group=self.group)
        pshard = self._pbuf[self.rank * self.chunk:(self.rank + 1) * self.chunk]
        for off, n, lr in self._pieces:
            self._adam(self._gshard[off:off + n], pshard[off:off + n], self.exp_avg[off:off + n],
                       self.exp_avg_sq[off:off + n], self.steps, lr, self.betas[0], self.betas[1], self.eps)
        dist.all_gather_into_tensor(self._pbuf, pshard, group=self.group)

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    def reset(self) -> None:
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.steps = 0

    def close(self) -> None:
        if self._registered:
            self._C.set_gradient_outputs(None)
            self._registered = False
