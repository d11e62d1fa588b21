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


def _worker_pipelined(rank, world, port, out_dict):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fgs_b200 import parallel
    from oracle import torch_oracle as O
    shapes = [(40, 3), (40, 3), (40, 4), (40, 1), (40, 1, 3), (40, 15, 3)]
    lrs = [1e-2, 2e-2, 3e-2, 4e-2, 5e-2, 6e-2]

    def fresh():
        torch.manual_seed(0)
        params = [torch.randn(s).requires_grad_(True) for s in shapes]
        for i, p in enumerate(params):
            p.grad = torch.randn(p.shape, generator=torch.Generator().manual_seed(100 * rank + i))
        opt = torch.optim.Adam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs)], lr=0.0, eps=1e-15)
        state = {id(p): (torch.zeros_like(p), torch.zeros_like(p)) for p in params}
        stepped = []

        def step_group(g):
            p = g['params'][0]
            m, v = state[id(p)]
            with torch.no_grad():
                O.adam_step(p, p.grad, m, v, 1, g['lr'])
            stepped.append(tuple(p.shape))
        return params, opt, step_group, stepped

    # pipelined: all-reduce largest first, step each tensor as its reduction completes
    pa, oa, sga, order = fresh()
    dens = torch.full((2, 40), float(rank + 1))
    parallel.allreduce_and_step(pa, oa, densification_info=dens, step_group=sga)
    # reference: all-reduce everything, then step every group
    pb, ob, sgb, _ = fresh()
    parallel.allreduce_gradients(pb)
    for g in ob.param_groups:
        sgb(g)
    out_dict[rank] = ([p.detach().clone() for p in pa], [p.detach().clone() for p in pb], order, dens.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_allreduce_and_step_equals_allreduce_then_step():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_pipelined, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        pa, pb, order, dens = out[rank]
        for a, b in zip(pa, pb):
            assert torch.equal(a, b), 'pipelined update differs from all-reduce followed by step'
        assert order[0] == (40, 15, 3) and len(order) == 6, order  # largest tensor first, every group stepped once
        assert torch.equal(dens, torch.full((2, 40), 3.0))
    for a, b in zip(out[0][0], out[1][0]):
        assert torch.equal(a, b), 'ranks diverged'


def test_allreduce_and_step_without_process_group_is_a_plain_step():
    from fgs_b200 import parallel
    p = torch.ones(3, requires_grad=True)
    q = torch.ones(2, requires_grad=True)
    p.grad, q.grad = torch.full((3,), 2.0), torch.full((2,), 1.0)
    opt = torch.optim.Adam([{'params': [p], 'lr': 0.1}, {'params': [q], 'lr': 0.1}], lr=0.0)
    seen = []
    parallel.allreduce_and_step([p], opt, step_group=lambda g: seen.append(g['params'][0].numel()))
    assert seen == [3, 2]  # the managed tensor, then the group the caller did not hand in


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


def _worker_flat(rank, world, port, out_dict):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fgs_b200 import parallel
    shapes = [(41, 3), (41, 3), (41, 4), (41, 1), (41, 1, 3), (41, 15, 3)]  # 41 * 3 = 123: regions need padding to float4s

    def grads():
        return [torch.randn(s, generator=torch.Generator().manual_seed(10 * rank + i)) for i, s in enumerate(shapes)]
    pa = [torch.zeros(s, requires_grad=True) for s in shapes]
    pb = [torch.zeros(s, requires_grad=True) for s in shapes]
    for p, g in zip(pa, grads()):
        p.grad = g
    for p, g in zip(pb, grads()):
        p.grad = g
    dens_a, dens_b = torch.full((2, 41), float(rank + 1)), torch.full((2, 41), float(rank + 1))
    flat = parallel.FlatGradients(pa, register=False)  # register=True hands the views to the CUDA backward: not on CPU
    flat.allreduce(pa, densification_info=dens_a)
    parallel.allreduce_gradients(pb, densification_info=dens_b)
    same_memory = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(pa, flat.views))
    padding_zero = all(float(flat.buffer[o + c:o + (c + 3) // 4 * 4].abs().sum()) == 0.0 for o, c in zip(flat.offsets, flat.counts))
    out_dict[rank] = ([p.grad.clone() for p in pa], [p.grad.clone() for p in pb], dens_a, dens_b, same_memory, padding_zero)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradients_single_allreduce_equals_six():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_flat, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        ga, gb, da, db, same_memory, padding_zero = out[rank]
        assert same_memory and padding_zero
        for a, b in zip(ga, gb):
            assert torch.equal(a, b), 'one flat all-reduce differs from six all-reduces'
        assert torch.equal(da, db) and float(da[0, 0]) == 3.0
    for a, b in zip(out[0][0], out[1][0]):
        assert torch.equal(a, b), 'ranks diverged'


def _worker_sharded(rank, world, port, out_dict):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fgs_b200 import parallel
    from oracle import torch_oracle as O
    shapes = [(37, 3), (37, 3), (37, 4), (37, 1), (37, 1, 3), (37, 15, 3)]  # 37: tensor regions with padding
    lrs = [1e-2, 2e-2, 3e-2, 4e-2, 5e-2, 6e-2]
    torch.manual_seed(0)
    init = [torch.randn(s) for s in shapes]

    def grad_of(r, step, i):
        return torch.randn(shapes[i], generator=torch.Generator().manual_seed(1000 * step + 10 * r + i))

    def oracle_adam(g, p, m, v, step, lr, b1, b2, eps):
        with torch.no_grad():
            O.adam_step(p, g, m, v, step, lr, b1, b2, eps)

    opt = parallel.ShardedAdam(init, lrs, adam_fn=oracle_adam, register=False)
    # yardstick: all-reduce (sum) + dense Adam on every rank
    dense = [t.clone() for t in init]
    state = [(torch.zeros_like(t), torch.zeros_like(t)) for t in init]
    for step in (1, 2, 3):
        for i, p in enumerate(opt.params):
            p.grad = grad_of(rank, step, i)
        opt.step()
        opt.zero_grad()
        for i, (p, (m, v)) in enumerate(zip(dense, state)):
            g = sum(grad_of(r, step, i) for r in range(world))
            O.adam_step(p, g, m, v, step, lrs[i])
    out_dict[rank] = ([p.detach().clone() for p in opt.params], dense)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_adam_equals_allreduce_then_dense_adam():
    """reduce-scatter -> Adam on 1/W of the flat buffer -> all-gather == all-reduce -> dense Adam (gloo, 2 ranks)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_sharded, args=(2, _free_port(), out), nprocs=2, join=True)
    for rank in range(2):
        got, want = out[rank]
        for a, b in zip(got, want):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    for a, b in zip(out[0][0], out[1][0]):
        assert torch.equal(a, b)  # replicas identical
