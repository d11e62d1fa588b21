"""Densification / pruning / reordering with the Adam state (reference Model.py:276-366,459-463).

CPU: the kernel SOURCE of csrc/densify.cu executed by the test-only CUDA emulator (tests/cuda_emu) against
oracle/densify_oracle.py -- adaptive density control (flags, order-preserving plan, fused row move with split
transform), prune, sort, extend, state reset, Morton codes.  GPU: the same through fgs_b200.densification on a real
FusedAdam optimizer, plus a 1 M-Gaussian run checked through size-independent properties."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import densify_oracle as DO

NAMES = DO.PARAMS
ROLE = {'means': 2, 'scales': 3}
ADC = dict(grad_threshold=0.0002, min_opacity=0.005, percent_dense=0.01, extent=4.0)


def make_state(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    p = {
        'means': r(n, 3) * 1.5,
        # raw log-scales around the "small" threshold log(0.01 * 4) = -3.2, a few above log(0.1 * 4) = -0.9
        'scales': r(n, 3) * 0.9 - 3.3,
        'rotations': r(n, 4),
        'opacities': r(n, 1) * 3.0,  # logit(0.005) = -5.3: a few percent are pruned
        'sh_coefficients_0': r(n, 1, 3),
        'sh_coefficients_rest': r(n, 15, 3) * 0.1,
    }
    p['rotations'][::97] *= 1e-6  # degenerate quaternions
    state = {k: v.float().contiguous() for k, v in p.items()}
    for k in NAMES:
        state['exp_avg.' + k] = (r(*state[k].shape) * 1e-3).float()
        state['exp_avg_sq.' + k] = (r(*state[k].shape) ** 2 * 1e-6).float()
    info = torch.stack([torch.randint(0, 40, (n,), generator=g).float(),
                        torch.rand(n, generator=g) * 0.012])  # accum / count around the 2e-4 threshold
    return state, info.contiguous()


def fixed_randn(seed):
    def f(shape):
        return torch.randn(shape, generator=torch.Generator().manual_seed(seed))
    return f


def compare_states(got: dict, want: dict, what=''):
    assert set(got) == set(want)
    for k in want:
        a, b = torch.as_tensor(got[k]), want[k]
        assert tuple(a.shape) == tuple(b.shape), (what, k, tuple(a.shape), tuple(b.shape))
        if k in ('means', 'scales'):  # split children: exp / log / rotation in float32
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), (what, k, float((a - b).abs().max()))
        else:
            assert torch.equal(a, b), (what, k)


# ---- emulator ---------------------------------------------------------------------------------------------------
def _emu():
    from tests import cuda_emu
    from fgs_b200._capi import RowTensor
    lib = cuda_emu.build('densify')
    P = C.c_void_p
    lib.fgs_densify_scratch_bytes.restype = C.c_size_t
    lib.fgs_densify_scratch_bytes.argtypes = [C.c_int32]
    lib.fgs_gather_rows.argtypes = [C.POINTER(RowTensor), C.c_int32, P, C.c_int64, C.c_int64, P, P, P, P]
    lib.fgs_zero_rows.argtypes = [P, P, C.c_int32, P, C.c_int64, P]
    lib.fgs_plan_count.argtypes = [P, C.c_int32, P, C.POINTER(C.c_uint32), P]
    lib.fgs_plan_emit.argtypes = [P, C.c_int32, P, P, P]
    lib.fgs_adc_flags.argtypes = [P, P, P, P, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_float, P, P]
    lib.fgs_morton_keys.argtypes = [P, C.c_int32, P, P, P, P, P, P]
    lib.fgs_gather_u32.argtypes = [P, P, P, P, C.c_int32, P]
    return lib, RowTensor


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def emu_move_rows(lib, RowTensor, state, plan, n_out, n_copy=None, appended=None, split=None):
    n_copy = n_out if n_copy is None else n_copy
    src = {k: np.ascontiguousarray(v.numpy()) for k, v in state.items()}
    dst = {k: np.full((n_out,) + v.shape[1:], np.nan, dtype=np.float32) for k, v in src.items()}
    app = {k: np.ascontiguousarray(v.numpy()) for k, v in (appended or {}).items()}
    rows = []
    for k in src:
        role = 1 if k.startswith('exp_avg') else ROLE.get(k, 0)
        rows.append(RowTensor(_ptr(src[k]), _ptr(app[k]) if k in app else None, dst[k].ctypes.data,
                              int(np.prod(src[k].shape[1:])), role))
    arr = (RowTensor * len(rows))(*rows)
    sp = [None, None, None] if split is None else [_ptr(a) for a in split]
    rc = lib.fgs_gather_rows(arr, len(rows), _ptr(plan), n_out, n_copy, sp[0], sp[1], sp[2], None)
    assert rc == 0
    return {k: torch.from_numpy(v) for k, v in dst.items()}


def emu_plan(lib, flags):
    n = flags.shape[0]
    scratch = np.zeros(lib.fgs_densify_scratch_bytes(n), dtype=np.uint8)
    totals = (C.c_uint32 * 4)()
    assert lib.fgs_plan_count(flags.ctypes.data, n, scratch.ctypes.data, totals, None) == 0
    kept, clones, pairs, splits = (int(v) for v in totals)
    plan = np.full(kept + clones + 2 * pairs, -1, dtype=np.int64)
    if plan.size:
        assert lib.fgs_plan_emit(flags.ctypes.data, n, scratch.ctypes.data, plan.ctypes.data, None) == 0
    return plan, (kept, clones, pairs, splits)


@pytest.mark.parametrize('n,prune_large', [(2500, True), (1024, False), (77, True)])
def test_adc_emulated_vs_oracle(n, prune_large):
    lib, RowTensor = _emu()
    state, info = make_state(n, seed=n)
    want, z = DO.adaptive_density_control(state, info, ADC['grad_threshold'], ADC['min_opacity'], prune_large,
                                          ADC['percent_dense'], ADC['extent'], fixed_randn(5))
    flags = np.zeros(n, dtype=np.uint8)
    rc = lib.fgs_adc_flags(info.numpy().ctypes.data, state['scales'].numpy().ctypes.data, state['rotations'].numpy().ctypes.data,
                           state['opacities'].numpy().ctypes.data, n, ADC['grad_threshold'],
                           math.log(ADC['percent_dense'] * ADC['extent']),
                           math.log(ADC['min_opacity'] / (1 - ADC['min_opacity'])), int(prune_large),
                           math.log(0.1 * ADC['extent']), flags.ctypes.data, None)
    assert rc == 0
    plan, (kept, clones, pairs, splits) = emu_plan(lib, flags)
    assert 2 * splits == z.shape[0] and splits > 0 and clones > 0 and kept < n  # the scene exercises every branch
    zs = np.ascontiguousarray(z.numpy().astype(np.float32))
    got = emu_move_rows(lib, RowTensor, state, plan, plan.shape[0],
                        split=(state['scales'].numpy(), state['rotations'].numpy(), zs))
    compare_states(got, want, f'adc n={n}')
    # Adam state: moved with kept originals, zero for every new Gaussian
    assert float(got['exp_avg.means'][kept:].abs().max()) == 0.0 and float(got['exp_avg.means'][:kept].abs().max()) > 0.0


def test_prune_sort_extend_reset_emulated_vs_oracle():
    lib, RowTensor = _emu()
    n = 1500
    state, _ = make_state(n, seed=2)
    g = torch.Generator().manual_seed(9)
    # prune
    prune_mask = torch.rand(n, generator=g) < 0.3
    flags = (~prune_mask).numpy().astype(np.uint8)
    plan, (kept, clones, pairs, splits) = emu_plan(lib, flags)
    assert (clones, pairs, splits) == (0, 0, 0) and kept == int((~prune_mask).sum())
    assert np.array_equal(plan, torch.nonzero(~prune_mask).flatten().numpy())
    compare_states(emu_move_rows(lib, RowTensor, state, plan, kept), DO.prune(state, prune_mask), 'prune')
    # nothing kept / everything kept
    plan0, t0 = emu_plan(lib, np.zeros(n, dtype=np.uint8))
    assert plan0.size == 0 and t0 == (0, 0, 0, 0)
    plan1, t1 = emu_plan(lib, np.ones(n, dtype=np.uint8))
    assert np.array_equal(plan1, np.arange(n)) and t1 == (n, 0, 0, 0)
    # sort: a torch int64 permutation is a plan
    order = torch.randperm(n, generator=g)
    compare_states(emu_move_rows(lib, RowTensor, state, order.numpy(), n), DO.sort(state, order), 'sort')
    # extend: 40 new rows, zero moments
    new = {k: torch.randn((40,) + tuple(state[k].shape[1:]), generator=g) for k in NAMES}
    compare_states(emu_move_rows(lib, RowTensor, state, None, n + 40, n_copy=n, appended=new), DO.extend(state, new), 'extend')
    # reset_state
    idx = torch.randperm(n, generator=g)[:123].to(torch.int64)
    mom = {k: v.clone().numpy() for k, v in state.items() if k.startswith('exp_avg')}
    ptrs = (C.c_void_p * len(mom))(*[a.ctypes.data for a in mom.values()])
    rows = (C.c_int32 * len(mom))(*[int(np.prod(a.shape[1:])) for a in mom.values()])
    assert lib.fgs_zero_rows(ptrs, rows, len(mom), idx.numpy().ctypes.data, idx.numel(), None) == 0
    want = DO.reset_state(state, idx)
    for k, a in mom.items():
        assert torch.equal(torch.from_numpy(a), want[k]), k


def test_morton_emulated_vs_oracle():
    lib, _ = _emu()
    n = 3000
    g = torch.Generator().manual_seed(4)
    means = (torch.randn(n, 3, generator=g) * torch.tensor([3.0, 0.5, 10.0])).float().contiguous()
    means[5] = means[6]  # equal codes: the order must be stable
    m = means.numpy()
    scratch = np.zeros(lib.fgs_densify_scratch_bytes(n), dtype=np.uint8)
    codes = np.zeros(n, dtype=np.int64)
    lo, hi, iota = (np.zeros(n, dtype=np.uint32) for _ in range(3))
    assert lib.fgs_morton_keys(m.ctypes.data, n, scratch.ctypes.data, codes.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                               iota.ctypes.data, None) == 0
    want = DO.morton_encode(means).numpy()
    assert np.array_equal(codes, want)
    assert np.array_equal((hi.astype(np.int64) << 32) | lo.astype(np.int64), want) and np.array_equal(iota, np.arange(n))
    assert int(want.max()) < 2 ** 63 and int(hi.max()) < 2 ** 31
    # degenerate bounding box (one axis constant)
    flat = m.copy()
    flat[:, 1] = 0.25
    assert lib.fgs_morton_keys(flat.ctypes.data, n, scratch.ctypes.data, codes.ctypes.data, None, None, None, None) == 0
    assert np.array_equal(codes, DO.morton_encode(torch.from_numpy(flat)).numpy())
    # gather_u32 (the step between the two radix passes) and the index widening
    perm = torch.randperm(n, generator=g).numpy().astype(np.uint32)
    out = np.zeros(n, dtype=np.uint32)
    wide = np.zeros(n, dtype=np.int64)
    assert lib.fgs_gather_u32(hi.ctypes.data, perm.ctypes.data, out.ctypes.data, wide.ctypes.data, n, None) == 0
    assert np.array_equal(out, hi[perm]) and np.array_equal(wide, perm.astype(np.int64))


# ---- GPU ----------------------------------------------------------------------------------------------------------
def _optimizer(state, device):
    import fgs_b200
    params = {k: torch.nn.Parameter(state[k].to(device)) for k in NAMES}
    opt = fgs_b200.FusedAdam([{'params': [params[k]], 'lr': 1e-3, 'name': k} for k in NAMES], lr=0.0, eps=1e-15)
    for k in NAMES:
        opt.state[params[k]] = {'step': 7, 'exp_avg': state['exp_avg.' + k].to(device).contiguous(),
                                'exp_avg_sq': state['exp_avg_sq.' + k].to(device).contiguous()}
    return opt


def _state_of(opt, groups):
    out = {}
    for name, p in groups.items():
        out[name] = p.detach().cpu()
        st = opt.state[p]
        assert st['step'] == 7
        out['exp_avg.' + name] = st['exp_avg'].cpu()
        out['exp_avg_sq.' + name] = st['exp_avg_sq'].cpu()
    assert [g['params'][0] for g in opt.param_groups] == [groups[g['name']] for g in opt.param_groups]
    assert len(opt.state) == len(groups)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize('n,prune_large', [(2500, True), (100_000, False)])
def test_adc_gpu_vs_oracle(cuda_device, n, prune_large):
    from fgs_b200 import densification as D
    state, info = make_state(n, seed=n)
    want, z = DO.adaptive_density_control(state, info, ADC['grad_threshold'], ADC['min_opacity'], prune_large,
                                          ADC['percent_dense'], ADC['extent'], fixed_randn(5))
    opt = _optimizer(state, cuda_device)
    groups = D.adaptive_density_control(opt, info.to(cuda_device), ADC['grad_threshold'], ADC['min_opacity'], prune_large,
                                        ADC['percent_dense'], ADC['extent'], randn=lambda shape: z.reshape(shape))
    compare_states(_state_of(opt, groups), want, f'adc gpu n={n}')


@pytest.mark.gpu
def test_prune_sort_extend_reset_morton_gpu(cuda_device, n=50_000):
    from fgs_b200 import densification as D
    state, info = make_state(n, seed=3)
    g = torch.Generator().manual_seed(11)
    prune_mask = torch.rand(n, generator=g) < 0.3
    opt = _optimizer(state, cuda_device)
    groups, (info_new,) = D.prune_param_groups(opt, (~prune_mask).to(cuda_device), extras=[info.t().contiguous().to(cuda_device)])
    s1 = DO.prune(state, prune_mask)
    compare_states(_state_of(opt, groups), s1, 'prune gpu')
    assert torch.equal(info_new.cpu(), info.t()[~prune_mask])
    m = s1['means'].shape[0]
    order = torch.randperm(m, generator=g)
    groups = D.sort_param_groups(opt, order.to(cuda_device))
    s2 = DO.sort(s1, order)
    compare_states(_state_of(opt, groups), s2, 'sort gpu')
    new = {k: torch.randn((321,) + tuple(state[k].shape[1:]), generator=g) for k in NAMES}
    groups = D.extend_param_groups(opt, {k: v.to(cuda_device) for k, v in new.items()})
    s3 = DO.extend(s2, new)
    compare_states(_state_of(opt, groups), s3, 'extend gpu')
    idx = torch.randperm(m + 321, generator=g)[:1000]
    D.reset_state(opt, idx.to(cuda_device))
    compare_states(_state_of(opt, groups), DO.reset_state(s3, idx), 'reset gpu')
    # Morton codes and order
    means = groups['means'].detach()
    codes = D.morton_encode(means)
    want = DO.morton_encode(means.cpu())
    assert torch.equal(codes.cpu(), want)
    order = D.morton_order(means)
    assert torch.equal(order.cpu(), torch.sort(want, stable=True).indices)
    # the optimizer still steps after all of this
    for p in groups.values():
        p.grad = torch.ones_like(p)
    opt.step()
    torch.cuda.synchronize()
    assert opt.state[groups['means']]['step'] == 8


@pytest.mark.gpu
def test_densify_full_size_properties(cuda_device):
    """BASELINE.json's 1 M Gaussians: permutation round trip, prune == boolean indexing, Morton order sorted."""
    from fgs_b200 import densification as D
    from fgs_b200 import synthetic as syn
    n = 1_000_000
    sc = syn.make_scene(n, 0)
    names = ('means', 'scales', 'rotations', 'opacities', 'sh_coefficients_0', 'sh_coefficients_rest')
    state = dict(zip(names, [p.float().contiguous() for p in sc]))
    for k in names:
        state['exp_avg.' + k] = torch.full_like(state[k], 0.5)
        state['exp_avg_sq.' + k] = torch.full_like(state[k], 0.25)
    opt = _optimizer(state, cuda_device)
    means0 = state['means'].to(cuda_device)
    order = D.morton_order(means0)
    codes = D.morton_encode(means0)
    assert bool((codes[order][1:] >= codes[order][:-1]).all()) and torch.equal(torch.sort(order).values, torch.arange(n, device=cuda_device))
    groups = D.sort_param_groups(opt, order)
    assert torch.equal(groups['sh_coefficients_rest'].detach(), state['sh_coefficients_rest'].to(cuda_device)[order])
    inverse = torch.empty_like(order)
    inverse[order] = torch.arange(n, device=cuda_device)
    groups = D.sort_param_groups(opt, inverse)
    for k in names:
        assert torch.equal(groups[k].detach().cpu(), state[k]), k
    keep = (torch.rand(n, generator=torch.Generator().manual_seed(1)) < 0.7).to(cuda_device)
    groups = D.prune_param_groups(opt, keep)
    for k in names:
        assert torch.equal(groups[k].detach().cpu(), state[k][keep.cpu()]), k
    assert float(opt.state[groups['means']]['exp_avg'].min()) == 0.5
