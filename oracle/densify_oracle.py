"""CPU ORACLE (test infrastructure only -- never imported by the product path) of the densification / pruning /
reordering step.

Restates, with plain torch CPU ops on a dict of tensors, what /root/reference/Model.py does to the six parameter tensors
and -- through NeRFICG's Optim.adam_utils, which is NOT in the reference tree -- to their Adam state:

  * ``adaptive_density_control``  Model.py:312-366 (the statements are restated in the reference's order, including
    the evaluation of the prune mask on the EXTENDED set and the randn_like sample layout of Model.py:332);
  * ``prune`` / ``sort``          Model.py:276-310;
  * ``extend``                    Model.py:340-347 (new rows appended, zero exp_avg / exp_avg_sq for them -- the
    behaviour of the original 3DGS ``cat_tensors_to_optimizer`` that adam_utils.extend_param_groups wraps);
  * ``reset_state``               Model.py:406,449;
  * ``morton_encode``             Model.py:459-463 calls CudaUtils.MortonEncoding.morton_encode (not in the tree): standard
    construction restated -- 21 bits per axis of the position normalised to the bounding box, x most significant.

**Parity unpinned** for the parts that live in NeRFICG (adam_utils row semantics, Morton code, the normalisation inside
quaternion_to_rotation_matrix): no reference source or fixture exists for them; everything visible in Model.py is
restated line by line.
"""
from __future__ import annotations

import math

import torch

PARAMS = ('means', 'scales', 'rotations', 'opacities', 'sh_coefficients_0', 'sh_coefficients_rest')


def quaternion_to_rotation_matrix(q: torch.Tensor) -> torch.Tensor:
    """(w, x, y, z), normalised (Cameras.utils.quaternion_to_rotation_matrix with its default normalize=True)."""
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)


def _select(state: dict, index_or_mask: torch.Tensor) -> dict:
    """prune_param_groups / sort_param_groups: every parameter and both moments take the same rows."""
    return {k: v[index_or_mask].contiguous() for k, v in state.items()}


def prune(state: dict, prune_mask: torch.Tensor) -> dict:
    return _select(state, ~prune_mask)


def sort(state: dict, ordering: torch.Tensor) -> dict:
    return _select(state, ordering)


def extend(state: dict, new: dict) -> dict:
    """extend_param_groups: rows appended; the moments ('exp_avg.<name>', 'exp_avg_sq.<name>') get zero rows."""
    out = {}
    for k, v in state.items():
        if k in new:
            out[k] = torch.cat([v, new[k]])
        else:
            name = k.split('.', 1)[1]
            out[k] = torch.cat([v, torch.zeros_like(new[name])])
    return out


def reset_state(state: dict, indices: torch.Tensor) -> dict:
    out = dict(state)
    for k, v in state.items():
        if k.startswith('exp_avg'):
            v = v.clone()
            v[indices] = 0
            out[k] = v
    return out


def adaptive_density_control(state: dict, densification_info: torch.Tensor, grad_threshold: float, min_opacity: float,
                             prune_large_gaussians: bool, percent_dense: float, extent: float, randn: torch.Tensor):
    """Model.py:312-366.  `state`: the six parameters plus 'exp_avg.<name>' / 'exp_avg_sq.<name>'.  `randn`: a
    generator of the (2 * n_split, 3) standard-normal tensor (callable taking the shape), so that the caller can hand
    the same samples to the kernels.  Returns (new state, randn tensor used)."""
    p = state
    densification_mask = densification_info[1] >= grad_threshold * densification_info[0].clamp_min(1.0)
    is_small = torch.max(p['scales'], dim=1).values <= math.log(percent_dense * extent)
    duplicate_mask = densification_mask & is_small
    split_mask = densification_mask & ~is_small
    dup = {k: p[k][duplicate_mask] for k in PARAMS}
    split_scales = p['scales'][split_mask].exp().expand(2, -1, -1).flatten(end_dim=1)
    split_rotations = p['rotations'][split_mask].expand(2, -1, -1).flatten(end_dim=1)
    z = randn(tuple(split_scales.shape))
    offsets = (quaternion_to_rotation_matrix(split_rotations) @ (split_scales * z)[..., None])[..., 0] \
        if split_scales.shape[0] else torch.zeros_like(split_scales)
    split = {
        'means': p['means'][split_mask].expand(2, -1, -1).flatten(end_dim=1) + offsets,
        'scales': split_scales.mul(0.625).log(),
        'rotations': split_rotations,
        'opacities': p['opacities'][split_mask].expand(2, -1, -1).flatten(end_dim=1),
        'sh_coefficients_0': p['sh_coefficients_0'][split_mask].expand(2, -1, -1, -1).flatten(end_dim=1),
        'sh_coefficients_rest': p['sh_coefficients_rest'][split_mask].expand(2, -1, -1, -1).flatten(end_dim=1),
    }
    n_new = dup['means'].shape[0] + split['means'].shape[0]
    ext = extend(state, {k: torch.cat([dup[k], split[k]]) for k in PARAMS})
    prune_mask = torch.cat([split_mask, torch.zeros(n_new, dtype=torch.bool)])
    prune_mask |= ext['opacities'].flatten() < math.log(min_opacity / (1 - min_opacity))
    prune_mask |= ext['rotations'].mul(ext['rotations']).sum(dim=1) < 1e-8
    if prune_large_gaussians:
        prune_mask |= ext['scales'].max(dim=1).values > math.log(0.1 * extent)
    return prune(ext, prune_mask), z


def morton_encode(means: torch.Tensor) -> torch.Tensor:
    """63-bit Morton code (int64): 21 bits per axis, x most significant, position normalised to the bounding box."""
    lo, hi = means.min(dim=0).values, means.max(dim=0).values
    extent = hi - lo
    u = torch.where(extent > 0, (means - lo) / extent, torch.zeros_like(means))
    q = (u * 2097151.0).clamp(0.0, 2097151.0).to(torch.int64)  # float32 arithmetic, truncation
    code = torch.zeros(means.shape[0], dtype=torch.int64)
    for bit in range(21):
        for c in range(3):
            code |= ((q[:, c] >> bit) & 1) << (3 * bit + (2 - c))
    return code


def mcmc_densification(state: dict, min_opacity: float, cap_max: int, multinomial):
    """Model.py:368-457 on the dict of tensors; `multinomial(weights, n)` is the sampler (shared with the code under
    test so that both draw the same indices).  relocation_adjustment is oracle/aux_oracle.py's restatement."""
    from oracle import aux_oracle as A
    st = {k: v.clone() for k, v in state.items()}
    logit_min = math.log(min_opacity / (1 - min_opacity))
    eps = torch.finfo(torch.float32).eps

    def adjusted(sampled):
        opacities = torch.sigmoid(st['opacities']).flatten()
        _, inverse, counts_per_unique = sampled.unique(sorted=False, return_inverse=True, return_counts=True)
        counts = counts_per_unique[inverse] + 1
        new_o, new_s = A.relocation_adjustment(opacities[sampled], st['scales'][sampled].exp(), counts)
        return new_o.reshape(-1, 1).clamp(min_opacity, 1.0 - eps).logit(), new_s.log()

    dead = st['opacities'].flatten() <= logit_min
    dead |= st['rotations'].mul(st['rotations']).sum(dim=1) < 1e-8
    n_dead = int(dead.sum())
    if n_dead > 0:
        dead_idx, alive_idx = torch.where(dead)[0], torch.where(~dead)[0]
        opacities = torch.sigmoid(st['opacities']).flatten()
        sampled = alive_idx[multinomial(opacities[alive_idx], n_dead)]
        adj_o, adj_s = adjusted(sampled)
        st['opacities'][sampled] = adj_o
        st['scales'][sampled] = adj_s
        for name in ('means', 'sh_coefficients_0', 'sh_coefficients_rest', 'rotations'):
            st[name][dead_idx] = st[name][sampled]
        st['opacities'][dead_idx] = adj_o
        st['scales'][dead_idx] = adj_s
        st = reset_state(st, sampled)
    n = st['means'].shape[0]
    n_added = max(0, min(int(cap_max), int(1.05 * n)) - n)
    if n_added > 0:
        opacities = torch.sigmoid(st['opacities']).flatten()
        sampled = multinomial(opacities, n_added)
        adj_o, adj_s = adjusted(sampled)
        st['opacities'][sampled] = adj_o
        st['scales'][sampled] = adj_s
        new = {name: st[name][sampled] for name in ('means', 'sh_coefficients_0', 'sh_coefficients_rest', 'rotations')}
        new['opacities'], new['scales'] = adj_o, adj_s
        st = extend(st, new)
        st = reset_state(st, sampled)
    return st
