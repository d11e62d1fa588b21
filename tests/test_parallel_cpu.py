"""World-size-2 gloo tests of the view-parallel host logic (fgs_b200/parallel.py) on CPU: sharding,
gradient all-reduce (sum semantics == gradient accumulation), identical optimizer step on every rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, out_dict):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fgs_b200 import parallel
    from oracle import torch_oracle as O
    torch.manual_seed(0)
    shapes = [(50, 3), (50, 3), (50, 4), (50, 1), (50, 1, 3), (50, 15, 3)]
    params = [torch.randn(s).requires_grad_(True) for s in shapes]

    def view_grad(v, shape):  # deterministic stand-in for one view's backward
        g = torch.Generator().manual_seed(1000 + v)
        return torch.randn(shape, generator=g)

    def render_and_backward(v):
        for p in params:
            g = view_grad(v, p.shape)
            p.grad = g if p.grad is None else p.grad + g

    dens = torch.zeros(2, 50)
    mine = parallel.shard_views(n_views, rank, world)
    n = parallel.accumulate_views(render_and_backward, mine)
    dens[0] += float(n)
    parallel.allreduce_gradients(params, densification_info=dens)
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in params]
    with torch.no_grad():
        for p, (m, v) in zip(params, state):
            O.adam_step(p, p.grad, m, v, 1, 1e-2)
    out_dict[rank] = ([p.detach().clone() for p in params], [p.grad.clone() for p in params], dens.clone(), mine)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_views_partition():
    from fgs_b200 import parallel
    for world in (1, 2, 3, 8):
        seen = sorted(v for r in range(world) for v in parallel.shard_views(8, r, world))
        assert seen == list(range(8))
    assert parallel.shard_views(8, 1, 2) == [1, 3, 5, 7]
    assert parallel.shard_views(3, 5, 8) == []


def test_two_rank_allreduce_equals_gradient_accumulation():
    world, n_views = 2, 8
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_views, out), nprocs=world, join=True)
    p0, g0, d0, v0 = out[0]
    p1, g1, d1, v1 = out[1]
    assert sorted(v0 + v1) == list(range(n_views))
    shapes = [(50, 3), (50, 3), (50, 4), (50, 1), (50, 1, 3), (50, 15, 3)]
    for a, b, ga, shape in zip(p0, p1, g0, shapes):
        assert torch.equal(a, b), 'ranks diverged after the optimizer step'
        ref = sum(torch.randn(shape, generator=torch.Generator().manual_seed(1000 + v)) for v in range(n_views))
        assert torch.allclose(ga, ref, atol=1e-5), 'all-reduced gradient != accumulated gradient of the 8 views'
    assert torch.equal(d0, d1) and float(d0[0, 0]) == float(n_views)


def test_allreduce_is_noop_without_process_group():
    from fgs_b200 import parallel
    p = torch.ones(3, requires_grad=True)
    p.grad = torch.full((3,), 2.0)
    parallel.allreduce_gradients([p])
    assert torch.equal(p.grad, torch.full((3,), 2.0))
