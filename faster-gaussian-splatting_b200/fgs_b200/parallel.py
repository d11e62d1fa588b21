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
