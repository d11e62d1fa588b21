"""PeerAdam (csrc/peer_adam.cu, fgs_b200/parallel.py): gradient reduction + Adam + parameter broadcast over NVLink
peer memory.  Host-side layout arithmetic on CPU; the kernel itself needs two GPUs of one node (skipped otherwise)."""
import ctypes as C
import os
import socket

import pytest
import torch

SHAPES = [(1003, 3), (1003, 3), (1003, 4), (1003, 1), (1003, 1, 3), (1003, 15, 3)]


def test_peer_layout_and_slices_partition_every_tensor():
    from fgs_b200 import _capi, parallel
    lib = _capi.load()
    offsets, counts, total = parallel.peer_layout(SHAPES)
    assert all(o % 4 == 0 for o in offsets) and total % 4 == 0
    assert counts == [1003 * 3, 1003 * 3, 1003 * 4, 1003, 1003 * 3, 1003 * 45]
    for a, b, c in zip(offsets, offsets[1:] + [total], counts):
        assert b - a == (c + 3) // 4 * 4                      # regions padded to float4s, back to back
    arr = lambda v: (C.c_uint64 * 6)(*v)
    for world in (1, 2, 3, 5, 8):
        shard = lib.fgs_peer_adam_shard_floats(arr(counts), world)
        covered = [0] * 6
        for rank in range(world):
            begin, count, shard_off = arr([0] * 6), arr([0] * 6), arr([0] * 6)
            assert lib.fgs_peer_adam_slice(arr(counts), world, rank, begin, count, shard_off) == 0
            for t in range(6):
                assert begin[t] == covered[t] and begin[t] % 4 == 0 and count[t] % 4 == 0
                covered[t] += count[t]
                assert shard_off[t] + count[t] <= shard
                if t:
                    assert shard_off[t] >= shard_off[t - 1]
        assert covered == [(c + 3) // 4 * 4 for c in counts]    # the W slices tile every padded tensor exactly
        assert shard * world >= sum(covered) and shard % 4 == 0
    assert lib.fgs_peer_adam_slice(arr(counts), 2, 2, arr([0] * 6), arr([0] * 6), arr([0] * 6)) != 0  # rank outside the world


def test_peer_adam_step_rejects_bad_arguments_without_touching_the_gpu():
    from fgs_b200 import _capi
    lib = _capi.load()
    arr64 = lambda v: (C.c_uint64 * 6)(*v)
    ptrs = (C.c_void_p * 2)(16, 32)
    rc = lib.fgs_peer_adam_step(ptrs, ptrs, 9, 0, 16, 16, arr64([0] * 6), arr64([4] * 6), (C.c_int32 * 6)(*[1] * 6),
                                (C.c_double * 6)(*[1e-3] * 6), 0.9, 0.999, 1e-15, None)
    assert rc != 0 and b'world' in lib.fgs_peer_last_error()
    rc = lib.fgs_peer_adam_step(ptrs, ptrs, 2, 0, 16, 16, arr64([2, 0, 0, 0, 0, 0]), arr64([4] * 6),
                                (C.c_int32 * 6)(*[1] * 6), (C.c_double * 6)(*[1e-3] * 6), 0.9, 0.999, 1e-15, None)
    assert rc != 0 and b'multiples of 4' in lib.fgs_peer_last_error()


@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_peer_adam_kernel_emulated_ranks(world):
    """The kernel SOURCE of csrc/peer_adam.cu under the test-only CUDA emulator (tests/cuda_emu): W "ranks" with their own
    gradient / parameter buffers in host memory run one after the other (every rank writes a disjoint slice of every
    peer's parameters, so sequential == concurrent), two steps; every replica must hold the parameters that
    sum-of-gradients + dense Adam produces, bit-identical across replicas, and the sharded moments must reassemble
    to the dense ones."""
    import numpy as np
    from fgs_b200 import parallel
    from tests import cuda_emu
    lib = cuda_emu.build('peer_adam')
    lib.fgs_peer_adam_shard_floats.restype = C.c_size_t
    shapes = [(101, 3), (101, 3), (101, 4), (101, 1), (101, 1, 3), (101, 15, 3)]
    offsets, counts, total = parallel.peer_layout(shapes)
    arr64 = lambda v: (C.c_uint64 * 6)(*v)
    rng = np.random.default_rng(world)
    lrs, b1, b2, eps = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4], 0.9, 0.999, 1e-15
    p0 = np.zeros(total, dtype=np.float32)
    for o, c in zip(offsets, counts):
        p0[o:o + c] = rng.standard_normal(c).astype(np.float32)
    params = [p0.copy() for _ in range(world)]
    shard = lib.fgs_peer_adam_shard_floats(arr64(counts), world)
    m_sh = [np.zeros(shard, dtype=np.float32) for _ in range(world)]
    v_sh = [np.zeros(shard, dtype=np.float32) for _ in range(world)]
    # dense reference in float32, gradients summed rank 0 first
    p_ref, m_ref, v_ref = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in (1, 2):
        grads = []
        for q in range(world):
            g = np.zeros(total, dtype=np.float32)
            for o, c in zip(offsets, counts):
                g[o:o + c] = (rng.standard_normal(c) * 1e-3).astype(np.float32)
            grads.append(g)
        gp = (C.c_void_p * world)(*[g.ctypes.data for g in grads])
        pp = (C.c_void_p * world)(*[p.ctypes.data for p in params])
        for rank in range(world):
            rc = lib.fgs_peer_adam_step(gp, pp, world, rank, C.c_void_p(m_sh[rank].ctypes.data), C.c_void_p(v_sh[rank].ctypes.data),
                                        arr64(offsets), arr64(counts), (C.c_int32 * 6)(*[step] * 6), (C.c_double * 6)(*lrs),
                                        C.c_double(b1), C.c_double(b2), C.c_double(eps), None)
            assert rc == 0, lib.fgs_peer_last_error()
        gsum = np.zeros(total, dtype=np.float32)
        for g in grads:
            gsum = gsum + g
        for t, (o, c) in enumerate(zip(offsets, counts)):
            sl = slice(o, o + c)
            g = gsum[sl]
            m_ref[sl] = np.float32(b1) * (m_ref[sl] - g) + g
            v_ref[sl] = np.float32(b2) * (v_ref[sl] - g * g) + g * g
            ss = np.float32(lrs[t] / (1.0 - b1 ** step))
            bb = np.float32(1.0 / np.sqrt(1.0 - b2 ** step))
            p_ref[sl] = p_ref[sl] - ss * m_ref[sl] / (np.sqrt(v_ref[sl]) * bb + np.float32(eps))
        for q in range(1, world):
            assert np.array_equal(params[q], params[0]), f'replica {q} differs after step {step}'
        for o, c, lr in zip(offsets, counts, lrs):
            # fma vs two roundings in the moments, amplified where beta1 (m - g) + g cancels: a fraction of one update
            np.testing.assert_allclose(params[0][o:o + c], p_ref[o:o + c], rtol=2e-6, atol=5e-7 + 1e-4 * lr)
    # the moment shards, reassembled through fgs_peer_adam_slice, are the dense moments
    m_full, v_full = np.zeros_like(p0), np.zeros_like(p0)
    for rank in range(world):
        begin, count, soff = arr64([0] * 6), arr64([0] * 6), arr64([0] * 6)
        assert lib.fgs_peer_adam_slice(arr64(counts), world, rank, begin, count, soff) == 0
        for t in range(6):
            lo = offsets[t] + begin[t]
            m_full[lo:lo + count[t]] = m_sh[rank][soff[t]:soff[t] + count[t]]
            v_full[lo:lo + count[t]] = v_sh[rank][soff[t]:soff[t] + count[t]]
    # aggregate comparison: fma vs two roundings where beta (x - g) + g cancels gives isolated 1e-4 relative outliers
    for o, c in zip(offsets, counts):
        for got, want in ((m_full, m_ref), (v_full, v_ref)):
            # v = beta2 (v - g^2) + g^2 loses three digits to cancellation (beta2 = 0.999), in the reference as here
            assert np.abs(got[o:o + c] - want[o:o + c]).sum() <= 2e-4 * np.abs(want[o:o + c]).sum()
        assert np.all(m_full[o + c:o + (c + 3) // 4 * 4] == 0)  # the padding of every tensor region stays zero


def _sh_basis_np(d, active):
    """Y_1..Y_15 of unit directions d (n,3), float32 like the kernels (sh_utils.cuh:32-69); zero beyond the active degree."""
    import numpy as np
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    f = np.float32
    Y = np.zeros((d.shape[0], 15), dtype=np.float32)
    if active > 1:
        Y[:, 0], Y[:, 1], Y[:, 2] = -f(0.48860251190291987) * y, f(0.48860251190291987) * z, -f(0.48860251190291987) * x
    if active > 4:
        xx, yy, zz, xy, xz, yz = x * x, y * y, z * z, x * y, x * z, y * z
        Y[:, 3] = f(1.0925484305920792) * xy
        Y[:, 4] = -f(1.0925484305920792) * yz
        Y[:, 5] = f(0.94617469575755997) * zz - f(0.31539156525251999)
        Y[:, 6] = -f(1.0925484305920792) * xz
        Y[:, 7] = f(0.54627421529603959) * (xx - yy)
    if active > 9:
        Y[:, 8] = y * (f(0.59004358992664352) * yy - f(1.7701307697799304) * xx)
        Y[:, 9] = f(2.8906114426405538) * xy * z
        Y[:, 10] = y * (f(0.45704579946446572) - f(2.2852289973223288) * zz)
        Y[:, 11] = z * (f(1.865881662950577) * zz - f(1.1195289977703462))
        Y[:, 12] = x * (f(0.45704579946446572) - f(2.2852289973223288) * zz)
        Y[:, 13] = f(1.4453057213202769) * z * (xx - yy)
        Y[:, 14] = x * (f(1.7701307697799304) * yy - f(0.59004358992664352) * xx)
    return Y


@pytest.mark.parametrize('world,n,active,broadcast,split', [(1, 70, 16, 1, False), (2, 101, 16, 1, False), (3, 333, 9, 1, False),
                                                            (8, 1000, 16, 0, False), (2, 64, 4, 0, False),
                                                            (2, 101, 16, 1, True), (3, 333, 9, 1, True), (8, 1000, 16, 1, True)])
def test_peer_adam_rows_kernel_emulated_ranks(world, n, active, broadcast, split):
    """fgs_peer_adam_rows_step (compact exchange) under the emulator: W ranks, every rank's gradient buffer holds what its
    backward would have written for its view -- the five small gradients, NO sh_rest gradient (poisoned with NaN: it must
    never be read) and the view's camera position in the slot behind the tensors.  After two steps every replica equals
    dense Adam on the SUMMED full gradients, where view q's sh_rest gradient is Y_k(dir_q) * dcol_q.
    `split`: the product's schedule -- fgs_peer_adam_small_step (five small tensors, float4 kernel, snapshot of the old means)
    followed by fgs_peer_adam_rows_step in sh_rest-only mode."""
    import numpy as np
    from fgs_b200 import parallel
    from tests import cuda_emu
    lib = cuda_emu.build('peer_adam')
    lib.fgs_peer_adam_rows_per_rank.restype = C.c_uint64
    lib.fgs_peer_adam_rows_per_rank.argtypes = [C.c_uint64, C.c_int32]
    shapes = [(n, 3), (n, 3), (n, 4), (n, 1), (n, 1, 3), (n, 15, 3)]
    kf = [3, 3, 4, 1, 3, 45]
    offsets, counts, total = parallel.peer_layout(shapes)
    cam_off = total
    arr64 = lambda v: (C.c_uint64 * 6)(*v)
    rng = np.random.default_rng(100 * world + n)
    lrs, b1, b2, eps = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4], 0.9, 0.999, 1e-15
    C0 = np.float32(0.28209479177387814)
    p0 = np.zeros(total, dtype=np.float32)
    for o, c in zip(offsets, counts):
        p0[o:o + c] = rng.standard_normal(c).astype(np.float32)
    params = [p0.copy() for _ in range(world)]
    per = int(lib.fgs_peer_adam_rows_per_rank(n, world))
    assert per % 32 == 0 and per * world >= n
    m_sh = [np.zeros(per * 59, dtype=np.float32) for _ in range(world)]
    v_sh = [np.zeros(per * 59, dtype=np.float32) for _ in range(world)]
    p_ref, m_ref, v_ref = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in (1, 2):
        means = p_ref[offsets[0]:offsets[0] + 3 * n].reshape(n, 3).copy()  # the means the "forward" of this step saw
        grads, full = [], np.zeros(total, dtype=np.float32)
        for q in range(world):
            g = np.zeros(total + 4, dtype=np.float32)
            for t in range(4):
                g[offsets[t]:offsets[t] + counts[t]] = (rng.standard_normal(counts[t]) * 1e-3).astype(np.float32)
            dcol = (rng.standard_normal((n, 3)) * 1e-3).astype(np.float32)
            dcol[rng.random(n) < 0.3] = 0.0  # not visible in this view
            cam = (rng.standard_normal(3) * 5).astype(np.float32)
            g[offsets[4]:offsets[4] + 3 * n] = (C0 * dcol).reshape(-1)
            g[offsets[5]:offsets[5] + 45 * n] = np.nan  # never exchanged
            g[cam_off:cam_off + 3] = cam
            grads.append(g)
            # the full gradient of this view, as the unfused backward would have produced it
            fg = g[:total].copy()
            r = means - cam[None, :]
            d = (r / np.sqrt((r * r).sum(1, keepdims=True))).astype(np.float32)
            dc = (C0 * dcol) * (np.float32(1.0) / C0)
            Y = _sh_basis_np(d, active)
            fg[offsets[5]:offsets[5] + 45 * n] = (Y[:, :, None] * dc[:, None, :]).astype(np.float32).reshape(-1)
            full = full + fg
        gp = (C.c_void_p * world)(*[g.ctypes.data for g in grads])
        pp = (C.c_void_p * world)(*[p.ctypes.data for p in params])
        for rank in range(world):
            snap = None
            if split:
                snap = np.full(per * 3 + 4, np.nan, dtype=np.float32)
                rc = lib.fgs_peer_adam_small_step(gp, pp, world, rank, C.c_void_p(m_sh[rank].ctypes.data),
                                                  C.c_void_p(v_sh[rank].ctypes.data), arr64(offsets), C.c_uint64(n), 15,
                                                  (C.c_int32 * 6)(*[step] * 6), (C.c_double * 6)(*lrs), C.c_double(b1),
                                                  C.c_double(b2), C.c_double(eps), C.c_void_p(snap.ctypes.data), None, None, None)
                assert rc == 0, lib.fgs_peer_last_error()
            rc = lib.fgs_peer_adam_rows_step(gp, pp, world, rank, C.c_void_p(m_sh[rank].ctypes.data), C.c_void_p(v_sh[rank].ctypes.data),
                                             arr64(offsets), C.c_uint64(n), 15, active, C.c_uint64(cam_off), broadcast,
                                             (C.c_int32 * 6)(*[step] * 6), (C.c_double * 6)(*lrs), C.c_double(b1), C.c_double(b2),
                                             C.c_double(eps), None, None, None if snap is None else C.c_void_p(snap.ctypes.data), 0, None)
            assert rc == 0, lib.fgs_peer_last_error()
        if not broadcast:  # the caller's job (copy engines on a side stream in the product): hand every slice of sh_rest round
            for rank in range(world):
                lo, hi = min(rank * per, n), min(rank * per + per, n)
                a, b = offsets[5] + lo * 45, offsets[5] + hi * 45
                for q in range(world):
                    params[q][a:b] = params[rank][a:b]
        for t, (o, c) in enumerate(zip(offsets, counts)):
            sl = slice(o, o + c)
            g = full[sl]
            m_ref[sl] = np.float32(b1) * (m_ref[sl] - g) + g
            v_ref[sl] = np.float32(b2) * (v_ref[sl] - g * g) + g * g
            ss = np.float32(lrs[t] / (1.0 - b1 ** step))
            bb = np.float32(1.0 / np.sqrt(1.0 - b2 ** step))
            p_ref[sl] = p_ref[sl] - ss * m_ref[sl] / (np.sqrt(v_ref[sl]) * bb + np.float32(eps))
        for q in range(1, world):
            assert np.array_equal(params[q], params[0]), f'replica {q} differs after step {step}'
        for t, (o, c, lr) in enumerate(zip(offsets, counts, lrs)):
            got, want = params[0][o:o + c] - p0[o:o + c], p_ref[o:o + c] - p0[o:o + c]
            assert np.isfinite(got).all()
            # rsqrt / fma vs numpy's sqrt and separate roundings: compared in aggregate, like every gradient in this repo
            assert np.abs(got - want).sum() <= 2e-4 * np.abs(want).sum(), (t, step)
    # moments of the rows a rank owns sit tensor after tensor in its shard
    for rank in range(world):
        lo, hi = min(rank * per, n), min(rank * per + per, n)
        off = 0
        for t in range(6):
            want = m_ref[offsets[t] + lo * kf[t]:offsets[t] + hi * kf[t]]
            got = m_sh[rank][off:off + (hi - lo) * kf[t]]
            assert np.abs(got - want).sum() <= 2e-4 * np.abs(want).sum() + 1e-12
            off += per * kf[t]


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import fgs_b200
    from fgs_b200 import parallel, synthetic as syn
    from tests import util
    n, w, h = 4000, 160, 120
    sc = syn.make_scene(n, 0)
    lrs = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]
    cams = [syn.make_camera(k, w, h) for k in range(4)]
    gis = [syn.make_grad_image(w, h, seed=k).to(dev) for k in range(4)]

    def settings(k):
        return fgs_b200.RasterizerSettings(*util.settings_for(cams[k], device=dev))

    # reference semantics: all-reduce (sum) of the gradients, dense FusedAdam on every rank
    pa = [p.to(dev).clone().requires_grad_(True) for p in sc]
    oa = fgs_b200.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15)
    for it in range(2):
        k = 2 * it + rank
        fgs_b200.diff_rasterize(*pa, torch.empty(0), settings(k)).backward(gis[k])
        parallel.allreduce_gradients(pa)
        oa.step()
        oa.zero_grad()
    results = {}

    def run_variant(name, make):
        ob = make()
        for it in range(2):
            k = 2 * it + rank
            fgs_b200.diff_rasterize(*ob.params, torch.empty(0), settings(k)).backward(gis[k])
            ob.step()
            ob.zero_grad()
        ob.close()
        torch.cuda.synchronize()
        dist.barrier()
        errs = [util.rel_err(b.detach() - p0.to(dev), a.detach() - p0.to(dev)) for a, b, p0 in zip(pa, ob.params, sc)]
        flat = torch.cat([p.detach().reshape(-1) for p in ob.params])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        results[name] = (errs, bool(torch.equal(both[0], both[1])))

    # one kernel over peer memory: W peer loads / stores per element
    run_variant('peer', lambda: parallel.PeerAdam([p.to(dev) for p in sc], lrs, multicast=False))
    # the same with the reduction and the broadcast done by the NVSwitch (multimem), where the box has NVLS
    try:
        run_variant('peer-mc', lambda: parallel.PeerAdam([p.to(dev) for p in sc], lrs, multicast=True))
    except RuntimeError as e:
        if 'no multicast address' not in str(e):
            raise
        results['peer-mc'] = None
    # NCCL reduce-scatter + sharded Adam + all-gather
    run_variant('sharded', lambda: parallel.ShardedAdam([p.to(dev) for p in sc], lrs))
    # compact exchange (the owner rebuilds the peers' sh_rest gradients), everything broadcast by the kernel ...
    run_variant('peer-rows-sync', lambda: parallel.PeerAdam([p.to(dev) for p in sc], lrs, compact=True, overlap_sh_broadcast=False))
    # ... and with the sh_rest broadcast on the side stream under the next forward (deferred colour kernel waits for it)
    run_variant('peer-rows', lambda: parallel.PeerAdam([p.to(dev) for p in sc], lrs, compact=True, overlap_sh_broadcast=True))
    # the same with the small tensors through the NVSwitch (multimem.ld_reduce / multimem.st), where the box has NVLS
    try:
        run_variant('peer-rows-small-mc', lambda: parallel.PeerAdam([p.to(dev) for p in sc], lrs, compact=True,
                                                                    overlap_sh_broadcast=True, small_multicast=True))
    except RuntimeError as e:
        if 'no multicast address' not in str(e):
            raise
        results['peer-rows-small-mc'] = None
    # one row-sliced kernel for everything (NVLS where available), only the sh_rest broadcast overlapped; and over plain peer
    # loads / stores and copy engines
    run_variant('peer-rows-mono', lambda: parallel.PeerAdam([p.to(dev) for p in sc], lrs, compact=True, overlap_sh_broadcast=True,
                                                            split=False))
    run_variant('peer-rows-mono-p2p', lambda: parallel.PeerAdam([p.to(dev) for p in sc], lrs, compact=True, overlap_sh_broadcast=True,
                                                                split=False, multicast=False))
    out[rank] = results
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_adam_equals_allreduce_then_fused_adam_on_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs of one node')
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for rank in range(2):
        for name, res in out[rank].items():
            if res is None:
                print(f'{name}: not available on this box (no NVLS multicast)')
                continue
            errs, identical = res
            print(f'rank {rank} {name}: update rel-L1 vs all-reduce + FusedAdam {["%.1e" % e for e in errs]}, replicas identical {identical}')
            assert identical, f'{name}: replicas diverged'
            assert max(errs) <= 2e-3, (name, errs)   # Adam's first steps are sign-like in the gradient: aggregate comparison
