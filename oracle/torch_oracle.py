"""CPU oracle for the differentiable Gaussian rasterizer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is product code: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` fallback may
import it, and only as the checker / reported CPU baseline -- never as the thing shipped.

What it is: a plain-PyTorch (CPU, dense, autograd-differentiable) restatement of the
reference algorithm, written from the reference's kernels:

  preprocess ........ /root/reference/FasterGSCudaBackend/FasterGSCudaBackend/rasterization/include/kernels_forward.cuh:14-209
  quaternion -> R ... .../include/kernel_utils.cuh:11-30
  exact tile test ... .../include/kernel_utils.cuh:61-180
  SH -> RGB ......... .../include/sh_utils.cuh:7-69
  blend ............. .../include/kernels_forward.cuh:362-498
  constants ......... .../include/rasterization_config.h:8-60
  Adam .............. FasterGSCudaBackend/FasterGSCudaBackend/adam/src/adam.cu:10-34,54-56

Gradients come from autograd through the restated forward; the places where the reference's
hand-written backward deviates from the true derivative are restated explicitly (detached
anti-aliasing factor, kernels_backward.cuh:137-146; `min(q,0)` ignored in the derivative,
kernels_backward.cuh:438-453).

Parity pin: the reference repository holds no tests, golden vectors or fixtures (SURVEY.md §4,
§8c).  This oracle is pinned instead against outputs of the *compiled reference itself* run on
a B200 (tests/golden/*.npz, produced by tests/golden/make_golden.py through baseline/_ref).
If those fixtures are absent the oracle is "parity unpinned".

The blend is dense (every visible Gaussian x every pixel) which is mathematically equal to the
tiled blend: any pixel with alpha >= 1/255 lies inside the Gaussian's opacity-aware bounding
box and inside a tile that passes the exact tile test; only float
round-off at the thresholds can differ, hence tolerance-based comparisons (mean-L1 <= 1e-4).
"""
from __future__ import annotations

import math
from typing import NamedTuple

import torch

# rasterization_config.h:10-22,55-60
DILATION = 0.3
DILATION_AA = 0.1
MIN_COV2D_DET = 1e-6
MIN_ALPHA = 1.0 / 255.0
MIN_ALPHA_RCP = 255.0
T_THRESHOLD = 1e-4
TILE_W, TILE_H = 16, 12

# sh_utils.cuh:9-30
C0 = 0.28209479177387814
C1 = 0.48860251190291987
C2a, C2b, C2c, C2d = 1.0925484305920792, 0.94617469575755997, 0.31539156525251999, 0.54627421529603959
C3a, C3b, C3c, C3d = 0.59004358992664352, 1.7701307697799304, 2.8906114426405538, 0.45704579946446572
C3e, C3f, C3g, C3h = 2.2852289973223288, 1.865881662950577, 1.1195289977703462, 1.4453057213202769


class Settings(NamedTuple):
    """Same fields and order as RasterizerSettings (torch_bindings/rasterization.py:8-21)."""
    w2c: torch.Tensor
    cam_position: torch.Tensor
    bg_color: torch.Tensor
    active_sh_bases: int
    width: int
    height: int
    focal_x: float
    focal_y: float
    center_x: float
    center_y: float
    near_plane: float
    far_plane: float
    proper_antialiasing: bool


def sh_to_color(sh0, sh_rest, means, cam_position, active_sh_bases):
    """sh_utils.cuh:32-69 (unclamped colour)."""
    result = 0.5 + C0 * sh0[:, 0, :]
    if active_sh_bases > 1:
        d = means - cam_position[None, :]
        d = d / d.norm(dim=1, keepdim=True)
        x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
        c = sh_rest
        result = result - C1 * y * c[:, 0] + C1 * z * c[:, 1] - C1 * x * c[:, 2]
        if active_sh_bases > 4:
            xx, yy, zz = x * x, y * y, z * z
            xy, xz, yz = x * y, x * z, y * z
            result = (result + C2a * xy * c[:, 3] - C2a * yz * c[:, 4] + (C2b * zz - C2c) * c[:, 5]
                      - C2a * xz * c[:, 6] + C2d * (xx - yy) * c[:, 7])
            if active_sh_bases > 9:
                result = (result + y * (C3a * yy - C3b * xx) * c[:, 8] + C3c * xy * z * c[:, 9]
                          + y * (C3d - C3e * zz) * c[:, 10] + z * (C3f * zz - C3g) * c[:, 11]
                          + x * (C3d - C3e * zz) * c[:, 12] + C3h * z * (xx - yy) * c[:, 13]
                          + x * (C3b * yy - C3a * xx) * c[:, 14])
    return result


def quaternion_to_rotation(q):
    """kernel_utils.cuh:11-30: (r,x,y,z) order, normalisation folded in via 1/|q|^2."""
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    rx, ry, rz = r * x, r * y, r * z
    norm_sq = r * r + xx + yy + zz
    s = 1.0 / norm_sq
    R = torch.stack([
        1.0 - 2.0 * (yy + zz) * s, 2.0 * (xy - rz) * s, 2.0 * (xz + ry) * s,
        2.0 * (xy + rz) * s, 1.0 - 2.0 * (xx + zz) * s, 2.0 * (yz - rx) * s,
        2.0 * (xz - ry) * s, 2.0 * (yz + rx) * s, 1.0 - 2.0 * (xx + yy) * s,
    ], dim=1).reshape(-1, 3, 3)
    return R, norm_sq


def will_primitive_contribute(mean_shifted, conic, tile_x, tile_y, tau):
    """kernel_utils.cuh:71-114, vectorised; all arguments broadcastable tensors."""
    mx, my = mean_shifted
    A, B, C = conic
    rect_min_x = (tile_x * TILE_W).to(mx.dtype)
    rect_min_y = (tile_y * TILE_H).to(mx.dtype)
    rect_max_x = ((tile_x + 1) * TILE_W - 1).to(mx.dtype)
    rect_max_y = ((tile_y + 1) * TILE_H - 1).to(mx.dtype)
    x_min_diff = rect_min_x - mx
    x_left = (x_min_diff >= 0).to(mx.dtype)
    not_in_x = x_left + (mx > rect_max_x).to(mx.dtype)
    y_min_diff = rect_min_y - my
    y_above = (y_min_diff >= 0).to(mx.dtype)
    not_in_y = y_above + (my > rect_max_y).to(mx.dtype)
    inside = (not_in_x + not_in_y) == 0
    corner_x = rect_max_x + x_left * (rect_min_x - rect_max_x)
    corner_y = rect_max_y + y_above * (rect_min_y - rect_max_y)
    diff_x, diff_y = mx - corner_x, my - corner_y
    dx = torch.copysign(torch.full_like(mx, TILE_W - 1.0), x_min_diff)
    dy = torch.copysign(torch.full_like(my, TILE_H - 1.0), y_min_diff)
    tx = not_in_y * ((dx * A * diff_x + dx * B * diff_y) / (dx * A * dx)).clamp(0.0, 1.0)
    ty = not_in_x * ((dy * B * diff_x + dy * C * diff_y) / (dy * C * dy)).clamp(0.0, 1.0)
    px, py = corner_x + tx * dx, corner_y + ty * dy
    ex, ey = mx - px, my - py
    power = 0.5 * (A * ex * ex + C * ey * ey) + B * ex * ey
    return inside | (power <= tau)


@torch.no_grad()
def count_touched_tiles(mean2d, conic, tau, tile_bounds):
    """Exact per-Gaussian tile count, kernel_utils.cuh:117-180 (sequential + cooperative parts
    compute the same predicate over the same candidate set)."""
    tx0, tx1, ty0, ty1 = tile_bounds
    w, h = tx1 - tx0, ty1 - ty0
    n = torch.zeros_like(tx0)
    ms = (mean2d[:, 0] - 0.5, mean2d[:, 1] - 0.5)
    max_w, max_h = int(w.max().item()) if w.numel() else 0, int(h.max().item()) if h.numel() else 0
    for oy in range(max_h):
        row = h > oy
        if not bool(row.any()):
            break
        for ox in range(max_w):
            sel = (row & (w > ox)).nonzero(as_tuple=True)[0]
            if sel.numel() == 0:
                break
            ok = will_primitive_contribute(
                (ms[0][sel], ms[1][sel]), (conic[sel, 0], conic[sel, 1], conic[sel, 2]),
                tx0[sel] + ox, ty0[sel] + oy, tau[sel])
            n[sel] += ok.to(n.dtype)
    return n


class Projected(NamedTuple):
    visible: torch.Tensor       # (N,) bool: n_touched_tiles > 0
    depth: torch.Tensor         # (N,)
    mean2d: torch.Tensor        # (N,2)
    conic: torch.Tensor         # (N,3)  (A, B, C) = (c, -b, a)/det
    opacity: torch.Tensor       # (N,)   activated (and AA-scaled)
    color: torch.Tensor         # (N,3)  unclamped
    screen_bounds: torch.Tensor  # (N,4) int: xmin, xmax, ymin, ymax
    n_touched_tiles: torch.Tensor  # (N,) int


def cov3d_from_six(c6):
    """(N,6) = (xx, xy, xz, yy, yz, zz) -> (N,3,3) symmetric."""
    xx, xy, xz, yy, yz, zz = c6.unbind(1)
    return torch.stack([xx, xy, xz, xy, yy, yz, xz, yz, zz], dim=1).reshape(-1, 3, 3)


def preprocess(means, scales, rotations, opacities, sh0, sh_rest, s: Settings, cov3d=None) -> Projected:
    """kernels_forward.cuh:14-209 (P1-P12 in SURVEY.md §8.1).  `cov3d` (N,6): the 3D covariance is given directly
    (scales / rotations are ignored) -- the input mode of the time-sliced 4D Gaussians (oracle/fourd_oracle.py)."""
    dt = means.dtype
    W = s.w2c[:3, :3].to(dt)
    t = s.w2c[:3, 3].to(dt)
    cam = means @ W.T + t[None, :]
    depth = cam[:, 2]
    active = (depth >= s.near_plane) & (depth <= s.far_plane)

    opacity = torch.sigmoid(opacities[:, 0])
    active = active & (opacity >= MIN_ALPHA)

    if cov3d is None:
        variance = torch.exp(2.0 * scales)
        R, norm_sq = quaternion_to_rotation(rotations)
        active = active & (norm_sq >= 1e-8)
        cov3d = R @ torch.diag_embed(variance) @ R.transpose(1, 2)
    else:
        cov3d = cov3d_from_six(cov3d)

    safe_depth = torch.where(active, depth, torch.ones_like(depth))
    x = cam[:, 0] / safe_depth
    y = cam[:, 1] / safe_depth
    width, height = float(s.width), float(s.height)
    clip_l = (-0.15 * width - s.center_x) / s.focal_x
    clip_r = (1.15 * width - s.center_x) / s.focal_x
    clip_t = (-0.15 * height - s.center_y) / s.focal_y
    clip_b = (1.15 * height - s.center_y) / s.focal_y
    xc = x.clamp(clip_l, clip_r)
    yc = y.clamp(clip_t, clip_b)
    j11 = s.focal_x / safe_depth
    j13 = -j11 * xc
    j22 = s.focal_y / safe_depth
    j23 = -j22 * yc
    jw1 = j11[:, None] * W[0][None, :] + j13[:, None] * W[2][None, :]
    jw2 = j22[:, None] * W[1][None, :] + j23[:, None] * W[2][None, :]
    jwc1 = torch.einsum('ni,nij->nj', jw1, cov3d)
    jwc2 = torch.einsum('ni,nij->nj', jw2, cov3d)
    a_raw = (jwc1 * jw1).sum(1)
    b = (jwc1 * jw2).sum(1)
    c_raw = (jwc2 * jw2).sum(1)
    det_raw = a_raw * c_raw - b * b
    k = DILATION_AA if s.proper_antialiasing else DILATION
    a, c = a_raw + k, c_raw + k
    det = a * c - b * b
    active = active & (det >= MIN_COV2D_DET)
    safe_det = torch.where(active, det, torch.ones_like(det))
    conic = torch.stack([c / safe_det, -b / safe_det, a / safe_det], dim=1)
    if s.proper_antialiasing:
        # the coverage factor is detached from cov2d (rasterization_config.h:12, kernels_backward.cuh:137-146)
        factor = torch.sqrt(torch.clamp(det_raw / safe_det, min=0.0)).detach()
        opacity = opacity * factor
        active = active & (opacity >= MIN_ALPHA)

    mean2d = torch.stack([x * s.focal_x + s.center_x, y * s.focal_y + s.center_y], dim=1)

    with torch.no_grad():
        safe_op = torch.where(active, opacity, torch.ones_like(opacity))
        tau = torch.log(safe_op * MIN_ALPHA_RCP)
        ext_x = torch.clamp(torch.sqrt(a * 2.0 * tau) - 0.5, min=0.0)
        ext_y = torch.clamp(torch.sqrt(c * 2.0 * tau) - 0.5, min=0.0)
        grid_w = (s.width + TILE_W - 1) // TILE_W
        grid_h = (s.height + TILE_H - 1) // TILE_H
        pw, ph = grid_w * TILE_W, grid_h * TILE_H
        big = 1.0e9
        m2 = torch.nan_to_num(mean2d.detach(), nan=-big, posinf=big, neginf=-big).clamp(-big, big)
        ex = torch.nan_to_num(ext_x, nan=0.0, posinf=big)
        ey = torch.nan_to_num(ext_y, nan=0.0, posinf=big)
        xmin = torch.floor(m2[:, 0] - ex).clamp(0, pw).to(torch.int64)
        xmax = torch.ceil(m2[:, 0] + ex).clamp(0, pw).to(torch.int64)
        ymin = torch.floor(m2[:, 1] - ey).clamp(0, ph).to(torch.int64)
        ymax = torch.ceil(m2[:, 1] + ey).clamp(0, ph).to(torch.int64)
        tx0, tx1 = xmin // TILE_W, (xmax + TILE_W - 1) // TILE_W
        ty0, ty1 = ymin // TILE_H, (ymax + TILE_H - 1) // TILE_H
        n_max = (tx1 - tx0) * (ty1 - ty0)
        active = active & (n_max > 0)
        idx = active.nonzero(as_tuple=True)[0]
        n_touched = torch.zeros_like(n_max)
        if idx.numel() > 0:
            n_touched[idx] = count_touched_tiles(
                mean2d.detach()[idx], conic.detach()[idx], tau[idx],
                (tx0[idx], tx1[idx], ty0[idx], ty1[idx]))
        visible = active & (n_touched > 0)
        bounds = torch.stack([xmin, xmax, ymin, ymax], dim=1)

    color = sh_to_color(sh0, sh_rest, means, s.cam_position.to(dt), s.active_sh_bases)
    return Projected(visible, depth, mean2d, conic, opacity, color, bounds, n_touched)


class Binned(NamedTuple):
    tile_ranges: torch.Tensor   # (T,2) int64: [start, end) of every tile's list inside `instances`
    instances: torch.Tensor     # (I,) int64: Gaussian index per instance, tile-major, front-to-back inside a tile
    tile_ids: torch.Tensor      # (I,) int64: tile id (ty * grid_w + tx) per instance, ascending


@torch.no_grad()
def binning(proj: Projected, s: Settings) -> Binned:
    """Integer restatement of the binning stage: kernels_forward.cuh:224-328 (create_instances_cu: one
    (tile id, Gaussian index) pair per candidate tile of the Gaussian's tile rectangle, kernel_utils.cuh:61-68, that
    passes the exact tile test, generated in global depth order), the stable sort by tile id (forward.cu:195-202)
    and extract_instance_ranges_cu (:330-348).  Depth order = ascending raw float bits of the positive depth
    (:205, forward.cu:104-110); Gaussians with EXACTLY equal depth are ordered by index here (the reference leaves
    that order to its atomic slot allocation, :204-208)."""
    grid_w = (s.width + TILE_W - 1) // TILE_W
    grid_h = (s.height + TILE_H - 1) // TILE_H
    vis = proj.visible.nonzero(as_tuple=True)[0]
    depth_bits = proj.depth[vis].to(torch.float32).view(torch.int32).to(torch.int64)
    order = vis[torch.argsort(depth_bits, stable=True)]
    b = proj.screen_bounds[order]
    tx0, tx1 = b[:, 0] // TILE_W, (b[:, 1] + TILE_W - 1) // TILE_W
    ty0, ty1 = b[:, 2] // TILE_H, (b[:, 3] + TILE_H - 1) // TILE_H
    mean2d, conic = proj.mean2d[order], proj.conic[order]
    tau = torch.log(proj.opacity[order] * MIN_ALPHA_RCP)
    ms = (mean2d[:, 0] - 0.5, mean2d[:, 1] - 0.5)
    w, h = tx1 - tx0, ty1 - ty0
    rank = torch.arange(order.numel())  # position in depth order
    keys, ranks = [], []
    for oy in range(int(h.max().item()) if h.numel() else 0):
        for ox in range(int(w.max().item()) if w.numel() else 0):
            sel = ((h > oy) & (w > ox)).nonzero(as_tuple=True)[0]
            if sel.numel() == 0:
                continue
            ok = will_primitive_contribute((ms[0][sel], ms[1][sel]), (conic[sel, 0], conic[sel, 1], conic[sel, 2]),
                                           tx0[sel] + ox, ty0[sel] + oy, tau[sel])
            sel = sel[ok]
            keys.append((ty0[sel] + oy) * grid_w + tx0[sel] + ox)
            ranks.append(rank[sel])
    if keys:
        keys, ranks = torch.cat(keys), torch.cat(ranks)
        # stable sort by tile id of pairs generated in depth order == sort by (tile id, depth rank)
        perm = torch.argsort(keys * (order.numel() + 1) + ranks)
        keys, ranks = keys[perm], ranks[perm]
    else:
        keys, ranks = torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64)
    counts = torch.bincount(keys, minlength=grid_w * grid_h)
    ends = torch.cumsum(counts, 0)
    ranges = torch.stack([ends - counts, ends], dim=1)
    return Binned(ranges, order[ranks], keys)


class RenderOut(NamedTuple):
    image: torch.Tensor           # (3,H,W)
    proj: Projected
    final_transmittance: torch.Tensor  # (H,W)
    n_blended: torch.Tensor       # (H,W) number of Gaussians that contributed
    order: torch.Tensor           # depth-sorted indices of the visible Gaussians


def render(means, scales, rotations, opacities, sh0, sh_rest, s: Settings, pixel_chunk: int = 16384,
           window=None, cov3d=None) -> RenderOut:
    """Forward of diff_rasterize: preprocess + depth order + dense front-to-back blend
    (kernels_forward.cuh:362-498, B1-B8 in SURVEY.md §8.1).  Differentiable.

    `window=(x0, y0, x1, y1)` renders only that pixel rectangle (exactly: Gaussians whose
    opacity-aware bounding box misses the window have alpha < 1/255 on all of its pixels); the
    returned image is (3, y1-y0, x1-x0).  Used to time a bounded sample of a large workload."""
    proj = preprocess(means, scales, rotations, opacities, sh0, sh_rest, s, cov3d=cov3d)
    vis_mask = proj.visible
    if window is not None:
        wx0, wy0, wx1, wy1 = window
        b = proj.screen_bounds
        vis_mask = vis_mask & (b[:, 0] < wx1) & (b[:, 1] > wx0) & (b[:, 2] < wy1) & (b[:, 3] > wy0)
    vis = vis_mask.nonzero(as_tuple=True)[0]
    order = vis[torch.argsort(proj.depth.detach()[vis], stable=True)]
    dt = means.dtype
    H, Wd = s.height, s.width
    bg = s.bg_color.to(dt)
    m2 = proj.mean2d[order]
    con = proj.conic[order]
    op = proj.opacity[order]
    col = torch.clamp(proj.color[order], min=0.0)
    if window is None:
        ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(Wd, dtype=dt), indexing='ij')
    else:
        ys, xs = torch.meshgrid(torch.arange(window[1], window[3], dtype=dt),
                                torch.arange(window[0], window[2], dtype=dt), indexing='ij')
        H, Wd = ys.shape
    px = xs.reshape(-1) + 0.5
    py = ys.reshape(-1) + 0.5
    P = px.numel()
    img_chunks, T_chunks, n_chunks = [], [], []
    for p0 in range(0, P, pixel_chunk):
        cx, cy = px[p0:p0 + pixel_chunk], py[p0:p0 + pixel_chunk]
        dx = m2[:, 0:1] - cx[None, :]
        dy = m2[:, 1:2] - cy[None, :]
        q = -0.5 * (con[:, 0:1] * dx * dx + con[:, 2:3] * dy * dy) - con[:, 1:2] * dx * dy
        q = q + (torch.clamp(q, max=0.0) - q).detach()   # min(q,0), ignored in the derivative
        alpha = op[:, None] * torch.exp(q)
        alpha = alpha * (alpha.detach() >= MIN_ALPHA).to(dt)
        one_minus = 1.0 - alpha
        T_incl = torch.cumprod(one_minus, dim=0)
        T_excl = torch.cat([torch.ones_like(T_incl[:1]), T_incl[:-1]], dim=0)
        alive = (T_excl.detach() >= T_THRESHOLD).to(dt)  # "blend, then test T < 1e-4" (:467-477)
        w = T_excl * alpha * alive
        rgb = torch.einsum('kp,kc->cp', w, col)
        T_final = torch.prod(1.0 - alpha * alive, dim=0)
        img_chunks.append(rgb + T_final[None, :] * bg[:, None])
        T_chunks.append(T_final.detach())
        n_chunks.append((w.detach() > 0).sum(0))
    image = torch.cat(img_chunks, dim=1).reshape(3, H, Wd)
    return RenderOut(image, proj, torch.cat(T_chunks).reshape(H, Wd), torch.cat(n_chunks).reshape(H, Wd), order)


def render_and_grads(params, s: Settings, grad_image, dtype=torch.float64, densification_info=None):
    """Image + gradients w.r.t. the six parameter tensors for a given dL/dimage, i.e. what
    `_C.forward` + `_C.backward` return (rasterization_api.cu:13-178).  Optionally accumulates
    densification_info like kernels_backward.cuh:194-201."""
    leaves = [p.detach().to(dtype).clone().requires_grad_(True) for p in params]
    s = s._replace(w2c=s.w2c.to(dtype), cam_position=s.cam_position.to(dtype), bg_color=s.bg_color.to(dtype))
    out = render(*leaves, s)
    out.proj.mean2d.retain_grad()
    out.image.backward(grad_image.to(dtype))
    grads = [torch.zeros_like(l) if l.grad is None else l.grad for l in leaves]
    # rows of invisible Gaussians are exactly zero in the reference (kernels_backward.cuh:45)
    inv = ~out.proj.visible
    grads = [g.masked_fill(inv.reshape(-1, *([1] * (g.dim() - 1))), 0.0) for g in grads]
    if densification_info is not None:
        g2 = out.proj.mean2d.grad
        if g2 is None:
            g2 = torch.zeros_like(out.proj.mean2d)
        visf = out.proj.visible.to(densification_info.dtype)
        densification_info[0] += visf
        ndc = 0.5 * torch.stack([g2[:, 0] * s.width, g2[:, 1] * s.height], dim=1)
        densification_info[1] += (ndc.norm(dim=1) * out.proj.visible).to(densification_info.dtype)
    return out, grads


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """adam.cu:24-33,54-56 (== torch.optim.Adam(eps=1e-15)); in place, returns nothing."""
    bias1_rcp = 1.0 / (1.0 - beta1 ** step)
    bias2_sqrt_rcp = 1.0 / math.sqrt(1.0 - beta2 ** step)
    step_size = lr * bias1_rcp
    m = beta1 * (exp_avg - grad) + grad
    v = beta2 * (exp_avg_sq - grad * grad) + grad * grad
    denom = torch.sqrt(v) * bias2_sqrt_rcp + eps
    param -= step_size * m / denom
    exp_avg.copy_(m)
    exp_avg_sq.copy_(v)


def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 10.0 * math.log10(1.0 / max(mse, 1e-20))


def pruning_scores(means, scales, rotations, opacities, sh0, sh_rest, s: Settings, dtype=torch.float64):
    """Speedy-Splat pruning scores of one view (kernels_pruning_scores.cuh:348-508): for every
    contributing (Gaussian, pixel) pair, (opacity * dL/dalpha)^2 with dL/d(pixel colour) == 1, summed
    over the pixels.  dL/dalpha comes from autograd with alpha as the leaf.  Returns (N,) scores."""
    params = [p.detach().to(dtype) for p in (means, scales, rotations, opacities, sh0, sh_rest)]
    s = s._replace(w2c=s.w2c.to(dtype), cam_position=s.cam_position.to(dtype), bg_color=s.bg_color.to(dtype))
    with torch.no_grad():
        proj = preprocess(*params, s)
    vis = proj.visible.nonzero(as_tuple=True)[0]
    order = vis[torch.argsort(proj.depth[vis], stable=True)]
    H, Wd = s.height, s.width
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(Wd, dtype=dtype), indexing='ij')
    px, py = xs.reshape(-1) + 0.5, ys.reshape(-1) + 0.5
    m2, con, op = proj.mean2d[order], proj.conic[order], proj.opacity[order]
    col = torch.clamp(proj.color[order], min=0.0)
    dx = m2[:, 0:1] - px[None, :]
    dy = m2[:, 1:2] - py[None, :]
    q = torch.clamp(-0.5 * (con[:, 0:1] * dx * dx + con[:, 2:3] * dy * dy) - con[:, 1:2] * dx * dy, max=0.0)
    alpha0 = op[:, None] * torch.exp(q)
    alpha0 = alpha0 * (alpha0 >= MIN_ALPHA).to(dtype)
    alpha = alpha0.clone().requires_grad_(True)
    T_incl = torch.cumprod(1.0 - alpha, dim=0)
    T_excl = torch.cat([torch.ones_like(T_incl[:1]), T_incl[:-1]], dim=0)
    alive = (T_excl.detach() >= T_THRESHOLD).to(dtype)
    w = T_excl * alpha * alive
    image = torch.einsum('kp,kc->cp', w, col) + torch.prod(1.0 - alpha * alive, dim=0)[None, :] * s.bg_color[:, None]
    image.sum().backward()
    contrib = (alpha0 > 0) & (alive > 0)
    per_pair = (op[:, None] * alpha.grad) ** 2 * contrib.to(dtype)
    scores = torch.zeros(means.shape[0], dtype=dtype)
    scores[order] = per_pair.sum(dim=1)
    return scores
