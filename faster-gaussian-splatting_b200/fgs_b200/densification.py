"""Densification / pruning / reordering of the Gaussians together with their Adam state.

Mirrors what /root/reference/Model.py does every DENSIFICATION_INTERVAL iterations, with the same function names and
argument meaning as the NeRFICG helpers it calls (``Optim.adam_utils``: Model.py:18,273,278,294,341,406,437,449;
``CudaUtils.MortonEncoding.morton_encode``: Model.py:10,461), which are not in the reference tree:

    prune_param_groups(optimizer, valid_mask)      -> {name: Parameter}    Model.py:278
    sort_param_groups(optimizer, ordering)         -> {name: Parameter}    Model.py:296
    extend_param_groups(optimizer, {name: rows})   -> {name: Parameter}    Model.py:341,437
    reset_state(optimizer, indices=...)                                    Model.py:406,449
    morton_encode(means) -> int64 codes                                    Model.py:461
    morton_order(means)  -> int64 ordering (== torch.argsort(morton_encode(means)), stable)
    adaptive_density_control(optimizer, densification_info, ...)           Model.py:312-366, fused
    mcmc_densification(optimizer, min_opacity, cap_max)                    Model.py:368-457 (3DGS-MCMC relocation + growth)

The optimizer is a ``torch.optim.Adam`` with ONE tensor per param group and a ``'name'`` per group (Model.py:238-245);
state keys ``step`` / ``exp_avg`` / ``exp_avg_sq`` are preserved.  Each call moves every parameter tensor and both
moments of every group (plus optional per-Gaussian extras such as ``densification_info`` / ``filter_3d``) with ONE
kernel launch behind the C ABI (csrc/densify.cu); torch only allocates the outputs.  No torch fallback exists."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _capi

_lib = _capi.load()

_ROLE = {'means': 2, 'scales': 3}  # FGS_ROW_MEANS / FGS_ROW_SCALES; every other parameter is FGS_ROW_COPY
_COPY, _STATE = 0, 1


RowTensor = _capi.RowTensor


def _stream() -> int:
    s = _capi.current_stream()
    _capi.load().fgs_join_deferred(s)  # these operators move sh_rest and its moments: a pending deferred update comes first
    return s


def _check(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (_capi.on_device(t) and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"'{name}' must be a contiguous CUDA float tensor.")
    return t


def _groups(optimizer):
    out = []
    for group in optimizer.param_groups:
        if len(group['params']) != 1:
            raise RuntimeError('one tensor per param group expected (Model.py:238-245)')
        out.append((group.get('name', str(len(out))), group, group['params'][0]))
    return out


def _row_floats(t: torch.Tensor) -> int:
    return int(t.numel() // t.shape[0]) if t.shape[0] > 0 else int(math.prod(t.shape[1:]))


def _move_rows(optimizer, plan, n_out: int, n_copy: int | None = None, appended: dict | None = None, split=None,
               extras=()):
    """Builds the new parameter / moment tensors of every group with one fgs_gather_rows launch and swaps them into
    the optimizer.  `plan`: int64 device tensor of n_copy entries or None (identity); rows >= n_copy come from
    `appended[name]` (parameters) or are zero (moments).  `extras`: row-major (N, ...) float tensors moved along
    (kind-independent copy; zeros behind n_copy).  Returns ({name: Parameter}, [new extras])."""
    groups = _groups(optimizer)
    dev = groups[0][2].device
    n_copy = n_out if n_copy is None else n_copy
    rows, keep_alive, new_params, new_states, new_extras = [], [], {}, {}, []

    def add(src, role, appended_t=None):
        _check(src, 'tensor')
        dst = torch.empty((n_out,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
        app_ptr = None
        if appended_t is not None:
            appended_t = _check(appended_t.to(device=dev, dtype=torch.float32).contiguous(), 'appended rows')
            if tuple(appended_t.shape) != (n_out - n_copy,) + tuple(src.shape[1:]):
                raise RuntimeError(f'appended rows have shape {tuple(appended_t.shape)}')
            keep_alive.append(appended_t)
            app_ptr = appended_t.data_ptr() if appended_t.numel() else None
        rows.append(RowTensor(src.data_ptr() if src.numel() else None, app_ptr, dst.data_ptr(), _row_floats(src), role))
        return dst

    for name, group, param in groups:
        app = appended.get(name) if appended is not None else None
        if appended is not None and app is None:
            raise RuntimeError(f"extend: no rows for param group '{name}'")
        new_params[name] = add(param.data, _ROLE.get(name, _COPY), app)
        st = optimizer.state.get(param, None)
        if st is not None and 'exp_avg' in st:
            new_states[name] = {'step': st['step'], 'exp_avg': add(st['exp_avg'], _STATE),
                                'exp_avg_sq': add(st['exp_avg_sq'], _STATE)}
    for t in extras:
        new_extras.append(add(t, _COPY))
    if len(rows) > 24:
        raise RuntimeError('at most 24 tensors per call (FGS_MAX_ROW_TENSORS)')
    if n_out > 0:
        sp = [None, None, None]
        if split is not None:
            sp = [_check(t, 'split input').data_ptr() for t in split]
        with _capi.device_guard(dev):
            arr = (RowTensor * len(rows))(*rows)
            rc = _lib.fgs_gather_rows(arr, len(rows), plan.data_ptr() if plan is not None else None, n_out, n_copy, sp[0], sp[1],
                                      sp[2], _stream())
            _capi.check(rc, 'gather_rows')
    out = {}
    for name, group, param in groups:
        new_p = torch.nn.Parameter(new_params[name], requires_grad=param.requires_grad)
        optimizer.state.pop(param, None)
        group['params'][0] = new_p
        if name in new_states:
            optimizer.state[new_p] = new_states[name]
        out[name] = new_p
    return out, new_extras


def _plan_from_flags(flags: torch.Tensor, n: int):
    """flags (n,) uint8 on the device -> (plan int64 (n_out,), totals [kept, clones, split pairs kept, splits])."""
    dev = flags.device
    with _capi.device_guard(dev):
        scratch = torch.empty(_lib.fgs_densify_scratch_bytes(n), dtype=torch.uint8, device=dev)
        totals = (C.c_uint32 * 4)()
        _capi.check(_lib.fgs_plan_count(flags.data_ptr(), n, scratch.data_ptr(), totals, _stream()), 'plan_count')
        kept, clones, pairs, splits = (int(v) for v in totals)
        n_out = kept + clones + 2 * pairs
        plan = torch.empty(n_out, dtype=torch.int64, device=dev)
        if n_out > 0:
            _capi.check(_lib.fgs_plan_emit(flags.data_ptr(), n, scratch.data_ptr(), plan.data_ptr(), _stream()), 'plan_emit')
    return plan, (kept, clones, pairs, splits)


def prune_param_groups(optimizer, valid_mask: torch.Tensor, extras=()):
    """Keeps the rows where `valid_mask` is True, in order (Model.py:276-292).  With `extras` returns
    ({name: Parameter}, [extras restricted to the kept rows]), else the dict alone."""
    n = _groups(optimizer)[0][2].shape[0]
    flags = valid_mask.to(torch.bool).contiguous().view(torch.uint8)  # True == 1 == "keep the Gaussian"
    if flags.shape[0] != n:
        raise RuntimeError('valid_mask must have one entry per Gaussian')
    plan, (kept, _, _, _) = _plan_from_flags(flags, n)
    out, ex = _move_rows(optimizer, plan, kept, extras=extras)
    return (out, ex) if extras else out


def sort_param_groups(optimizer, ordering: torch.Tensor, extras=()):
    """Applies the permutation / index list `ordering` (int64) to every row (Model.py:294-310)."""
    ordering = ordering.to(torch.int64).contiguous()
    out, ex = _move_rows(optimizer, ordering, ordering.shape[0], extras=extras)
    return (out, ex) if extras else out


def extend_param_groups(optimizer, new_tensors: dict):
    """Appends the rows `new_tensors[name]` to every param group; their Adam state is zero (Model.py:340-347)."""
    groups = _groups(optimizer)
    n = groups[0][2].shape[0]
    n_new = next(iter(new_tensors.values())).shape[0]
    out, _ = _move_rows(optimizer, None, n + n_new, n_copy=n, appended=new_tensors)
    return out


def reset_state(optimizer, indices: torch.Tensor) -> None:
    """Zeroes exp_avg / exp_avg_sq of the rows `indices` in every group (Model.py:406,449)."""
    indices = indices.to(torch.int64).contiguous()
    if indices.numel() == 0:
        return
    tensors = []
    for _, _, param in _groups(optimizer):
        st = optimizer.state.get(param, None)
        if st is not None and 'exp_avg' in st:
            tensors += [_check(st['exp_avg'], 'exp_avg'), _check(st['exp_avg_sq'], 'exp_avg_sq')]
    if not tensors:
        return
    with _capi.device_guard(indices.device):
        ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        rows = (C.c_int32 * len(tensors))(*[_row_floats(t) for t in tensors])
        _capi.check(_lib.fgs_zero_rows(ptrs, rows, len(tensors), indices.data_ptr(), indices.numel(), _stream()), 'zero_rows')


def morton_encode(means: torch.Tensor) -> torch.Tensor:
    """63-bit Morton codes (int64) of the positions: 21 bits per axis of the position normalised to the bounding
    box, x most significant (stands in for CudaUtils.MortonEncoding.morton_encode, Model.py:461)."""
    _check(means, 'means')
    n = means.shape[0]
    codes = torch.empty(n, dtype=torch.int64, device=means.device)
    with _capi.device_guard(means.device):
        scratch = torch.empty(_lib.fgs_densify_scratch_bytes(n), dtype=torch.uint8, device=means.device)
        _capi.check(_lib.fgs_morton_keys(means.data_ptr(), n, scratch.data_ptr(), codes.data_ptr(), None, None, None, _stream()),
                    'morton_keys')
    return codes


def morton_order(means: torch.Tensor) -> torch.Tensor:
    """int64 ordering that sorts the Gaussians by Morton code (stable), i.e. torch.argsort(morton_encode(means),
    stable=True), computed with the library's own radix sort: two stable 32-bit passes over the 63-bit code."""
    _check(means, 'means')
    n, dev = means.shape[0], means.device
    order = torch.empty(n, dtype=torch.int64, device=dev)
    if n == 0:
        return order
    with _capi.device_guard(dev):
        u32 = lambda: torch.empty(n, dtype=torch.int32, device=dev)
        lo, hi, idx, lo_alt, idx_alt, hi_sorted = u32(), u32(), u32(), u32(), u32(), u32()
        scratch = torch.empty(_lib.fgs_densify_scratch_bytes(n), dtype=torch.uint8, device=dev)
        _capi.check(_lib.fgs_morton_keys(means.data_ptr(), n, scratch.data_ptr(), None, lo.data_ptr(), hi.data_ptr(), idx.data_ptr(),
                                         _stream()), 'morton_keys')
        n_dev = torch.tensor([n], dtype=torch.int32, device=dev)
        sb = _lib.fgs_sort_scratch_bytes(n, 32)
        sort_scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
        in_alt = C.c_int32(0)
        _capi.check(_lib.fgs_sort_pairs_u32(lo.data_ptr(), lo_alt.data_ptr(), idx.data_ptr(), idx_alt.data_ptr(), n_dev.data_ptr(), n,
                                            32, sort_scratch.data_ptr(), sb, C.byref(in_alt), _stream()), 'sort (low half)')
        idx1, idx1_alt = (idx_alt, idx) if in_alt.value else (idx, idx_alt)
        # the high halves in the order of the first pass; 31 significant bits (the code has 63)
        _capi.check(_lib.fgs_gather_u32(hi.data_ptr(), idx1.data_ptr(), hi_sorted.data_ptr(), None, n, _stream()), 'gather_u32')
        _capi.check(_lib.fgs_sort_pairs_u32(hi_sorted.data_ptr(), lo_alt.data_ptr(), idx1.data_ptr(), idx1_alt.data_ptr(),
                                            n_dev.data_ptr(), n, 31, sort_scratch.data_ptr(), sb, C.byref(in_alt), _stream()),
                    'sort (high half)')
        final = idx1_alt if in_alt.value else idx1
        _capi.check(_lib.fgs_gather_u32(None, final.data_ptr(), None, order.data_ptr(), n, _stream()), 'index widening')
    return order


def adaptive_density_control(optimizer, densification_info: torch.Tensor, grad_threshold: float, min_opacity: float,
                             prune_large_gaussians: bool, percent_dense: float, training_cameras_extent: float,
                             randn=None, extras=()):
    """Model.py:312-366 in three small kernels + one row-moving launch: clone the small / split the large Gaussians
    whose accumulated screen-space gradient exceeds the threshold, then prune (split parents, transparent, degenerate
    and -- optionally -- too large ones), all evaluated exactly as the reference does on the extended set, output in
    the reference's order.  `randn(shape)` supplies the (2 * n_split, 3) normal samples (default: torch.randn on the
    device, as Model.py:332).  Returns {name: Parameter} like the reference's param_groups dict (and the moved
    `extras`, if any: the reference drops densification_info / filter_3d here, Model.py:355-357)."""
    groups = {name: param for name, _, param in _groups(optimizer)}
    means, scales, rotations, opacities = (groups[k].data for k in ('means', 'scales', 'rotations', 'opacities'))
    n, dev = means.shape[0], means.device
    info = _check(densification_info, 'densification_info')
    if tuple(info.shape) != (2, n):
        raise RuntimeError("'densification_info' must have shape (2, N)")
    flags = torch.empty(n, dtype=torch.uint8, device=dev)
    with _capi.device_guard(dev):
        rc = _lib.fgs_adc_flags(info.data_ptr(), _check(scales, 'scales').data_ptr(), _check(rotations, 'rotations').data_ptr(),
                                _check(opacities, 'opacities').data_ptr(), n, float(grad_threshold),
                                float(math.log(percent_dense * training_cameras_extent)),
                                float(math.log(min_opacity / (1 - min_opacity))), int(bool(prune_large_gaussians)),
                                float(math.log(0.1 * training_cameras_extent)), flags.data_ptr(), _stream())
        _capi.check(rc, 'adc_flags')
    plan, (kept, clones, pairs, splits) = _plan_from_flags(flags, n)
    shape = (2 * splits, 3)
    z = (randn(shape) if randn is not None else torch.randn(shape, device=dev)).to(device=dev, dtype=torch.float32).contiguous()
    if tuple(z.shape) != shape:
        raise RuntimeError(f'randn must return a tensor of shape {shape}')
    out, ex = _move_rows(optimizer, plan, kept + clones + 2 * pairs, split=(scales, rotations, z), extras=extras)
    return (out, ex) if extras else out


def mcmc_densification(optimizer, min_opacity: float, cap_max: int, multinomial=None):
    """Model.py:368-457 (3DGS-MCMC): dead Gaussians (opacity <= min_opacity or a degenerate quaternion) are relocated
    onto samples of the live ones drawn with probability proportional to opacity, then the set grows by 5 % up to
    `cap_max` the same way; sampled Gaussians share their opacity / scale with their copies (relocation_adjustment,
    Eq. 9) and get a fresh Adam state.  The sampling and the index assignments are torch ops as in the reference; the
    adjustment, the state reset and the growth run through the C ABI (fgs_relocation_adjustment, fgs_zero_rows,
    fgs_gather_rows).  `multinomial(weights, n)` defaults to torch.multinomial(weights, n, replacement=True).
    Returns {name: Parameter}."""
    from . import _C
    if multinomial is None:
        multinomial = lambda weights, n: torch.multinomial(weights, n, replacement=True)
    logit_min = math.log(min_opacity / (1 - min_opacity))
    eps = torch.finfo(torch.float32).eps

    def adjusted(groups, sampled):  # Model.py:384-391
        opacities = torch.sigmoid(groups['opacities'].data).flatten()
        _, inverse, counts_per_unique = sampled.unique(sorted=False, return_inverse=True, return_counts=True)
        counts = counts_per_unique[inverse] + 1  # +1 for the original Gaussian
        new_o, new_s = _C.relocation_adjustment(opacities[sampled], groups['scales'].data[sampled].exp(), counts)
        return new_o.clamp(min_opacity, 1.0 - eps).logit(), new_s.log()

    groups = {name: param for name, _, param in _groups(optimizer)}
    with torch.no_grad():
        dead_mask = groups['opacities'].data.flatten() <= logit_min
        dead_mask |= groups['rotations'].data.mul(groups['rotations'].data).sum(dim=1) < 1e-8
        n_dead = int(dead_mask.sum().item())
        if n_dead > 0:
            dead_indices = torch.where(dead_mask)[0]
            alive_indices = torch.where(~dead_mask)[0]
            opacities = torch.sigmoid(groups['opacities'].data).flatten()
            sampled = alive_indices[multinomial(opacities[alive_indices], n_dead)]
            adj_o, adj_s = adjusted(groups, sampled)
            groups['opacities'].data[sampled] = adj_o
            groups['scales'].data[sampled] = adj_s
            for name in ('means', 'sh_coefficients_0', 'sh_coefficients_rest', 'rotations'):
                groups[name].data[dead_indices] = groups[name].data[sampled]
            groups['opacities'].data[dead_indices] = adj_o
            groups['scales'].data[dead_indices] = adj_s
            reset_state(optimizer, sampled)
        n = groups['means'].shape[0]
        n_added = max(0, min(int(cap_max), int(1.05 * n)) - n)
        if n_added > 0:
            opacities = torch.sigmoid(groups['opacities'].data).flatten()
            sampled = multinomial(opacities, n_added)
            adj_o, adj_s = adjusted(groups, sampled)
            groups['opacities'].data[sampled] = adj_o
            groups['scales'].data[sampled] = adj_s
            new = {name: groups[name].data[sampled] for name in ('means', 'sh_coefficients_0', 'sh_coefficients_rest', 'rotations')}
            new['opacities'], new['scales'] = adj_o, adj_s
            groups = extend_param_groups(optimizer, new)
            reset_state(optimizer, sampled)
    return groups
