"""`_C`: the native-module surface of the reference, served by libfgs_b200.so.

Same function names, argument order, argument meaning and return tuples as the reference's
pybind11 module (/root/reference/FasterGSCudaBackend/FasterGSCudaBackend/torch_bindings/bindings.cpp:12-21,
rasterization/include/rasterization_api.h:8-106, adam/include/adam.h:7-16), so that the reference's
own torch_bindings/*.py run unchanged on top of this module.  torch is used for device memory and
the current stream only; all computation happens in the hand-written sm_100a kernels behind the
C ABI (include/fgs_b200.h).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi

_lib = _capi.load()

# ---- scratch management: the C ABI grows caller-owned buffers through callbacks ----------------
# Every call that needs scratch opens its own context (device + the buffers handed out so far); the `user` word the
# library passes back to the callback is (context id << 2 | which buffer), so calls on different devices or from different
# host threads never see each other's buffers (ctypes releases the GIL while the library runs).
_PRIMITIVE, _TILE, _INSTANCE = 1, 2, 3
_contexts: dict[int, '_Scratch'] = {}
_next_context = [1]


class _Scratch:
    def __init__(self, device):
        self.device = device
        self.buffers: dict[int, torch.Tensor] = {}
        self.key = _next_context[0]
        _next_context[0] += 1

    def __enter__(self):
        _contexts[self.key] = self
        return self

    def __exit__(self, *exc):
        _contexts.pop(self.key, None)
        return False

    def user(self, which: int) -> int:
        return (self.key << 2) | which


def _resize(user, nbytes):
    ctx = _contexts[int(user) >> 2]
    which = int(user) & 3
    if int(nbytes) == 0 and which in ctx.buffers:  # query: the buffer handed out for this user last time
        return ctx.buffers[which].data_ptr()
    t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=ctx.device)
    ctx.buffers[which] = t
    return t.data_ptr()


_resize_cb = _capi.ResizeFn(_resize)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _check_param(t, name: str) -> None:
    # the reference only checks this in debug builds (rasterization_api.cu:35-41); here it is always on
    if t is None and name == 'rotations':  # covariance-input mode: `scales` holds (N,6) covariances (include/fgs_b200.h)
        return
    if not (_capi.on_device(t) and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"Input tensor '{name}' must be a contiguous CUDA float tensor.")


def _aux(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


def _settings(active_sh_bases, total_rest, width, height, fx, fy, cx, cy, near, far, aa) -> _capi.Settings:
    return _capi.Settings(int(active_sh_bases), int(total_rest), int(width), int(height), float(fx), float(fy),
                          float(cx), float(cy), float(near), float(far), int(bool(aa)))


def _stream() -> int:
    return _capi.current_stream()


def forward(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest, w2c, cam_position,
            bg_color, active_sh_bases, width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
            proper_antialiasing):
    """rasterization_api.h:8-28.  Returns (image, primitive_buffers, tile_buffers, instance_buffers,
    bucket_buffers, n_instances, n_buckets, selector); the buffers are opaque."""
    return forward_ex(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest, w2c, cam_position,
                      bg_color, active_sh_bases, width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
                      proper_antialiasing)


def forward_ex(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest, w2c, cam_position,
               bg_color, active_sh_bases, width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
               proper_antialiasing, defer_color=False, color_event=None):
    """`forward` with the extension of include/fgs_b200.h "deferred SH schedule": `defer_color` = geometry pass without
    SH, colours evaluated after the sorts -- a pending deferred sh_rest update and `color_event` (a torch.cuda.Event,
    optional) are only waited for there.  Same image, same return tuple."""
    for t, name in ((means, 'means'), (scales, 'scales'), (rotations, 'rotations'), (opacities, 'opacities'),
                    (sh_coefficients_0, 'sh_coefficients_0'), (sh_coefficients_rest, 'sh_coefficients_rest')):
        _check_param(t, name)
    dev = means.device
    with _capi.device_guard(dev):
        n = means.shape[0]
        total_rest = sh_coefficients_rest.shape[1] if sh_coefficients_rest.dim() > 1 else 0
        image = torch.empty((3, int(height), int(width)), dtype=torch.float32, device=dev)
        w2c_c, pos_c, bg_c = _aux(w2c, dev), _aux(cam_position, dev), _aux(bg_color, dev)
        st = _settings(active_sh_bases, total_rest, width, height, focal_x, focal_y, center_x, center_y, near_plane,
                       far_plane, proper_antialiasing)
        state = _capi.ForwardState()
        with _Scratch(dev) as ctx:
            args = (_resize_cb, ctx.user(_PRIMITIVE), ctx.user(_TILE), ctx.user(_INSTANCE), _ptr(means),
                    _ptr(scales), _ptr(rotations), _ptr(opacities), _ptr(sh_coefficients_0),
                    _ptr(sh_coefficients_rest), _ptr(w2c_c), _ptr(pos_c), _ptr(bg_c), _ptr(image), n,
                    C.byref(st), _stream(), C.byref(state))
            if defer_color:
                rc = _lib.fgs_forward_ex(*args, 1, None if color_event is None else color_event.cuda_event)
            else:
                rc = _lib.fgs_forward(*args)
            _capi.check(rc, 'forward')
        empty = torch.empty(0, dtype=torch.uint8, device=dev)
        prim, tile, inst = (ctx.buffers.get(k, empty) for k in (_PRIMITIVE, _TILE, _INSTANCE))
    return image, prim, tile, inst, empty, state.n_instances, state.n_buckets, state.selector


def _backward_common(image, primitive_buffers, w2c, cam_position, bg_color, grad_image, densification_info,
                     sh_coefficients_rest, active_sh_bases, width, height, focal_x, focal_y, center_x, center_y,
                     near_plane, far_plane, proper_antialiasing, n_instances):
    dev = image.device
    n_total_rest = sh_coefficients_rest.shape[1] if sh_coefficients_rest.dim() > 1 else 0
    st = _settings(active_sh_bases, n_total_rest, width, height, focal_x, focal_y, center_x, center_y, near_plane,
                   far_plane, proper_antialiasing)
    state = _capi.ForwardState(int(n_instances), 0, 0, 0)
    w2c_c, pos_c, bg_c = _aux(w2c, dev), _aux(cam_position, dev), _aux(bg_color, dev)
    gi = grad_image.to(dtype=torch.float32).contiguous()
    dens_ptr = None
    if densification_info is not None and densification_info.numel() > 0:
        if not (_capi.on_device(densification_info) and densification_info.dtype == torch.float32
                and densification_info.is_contiguous()):
            raise RuntimeError("'densification_info' must be a contiguous CUDA float tensor.")
        dens_ptr = _ptr(densification_info)
    return dev, st, state, w2c_c, pos_c, bg_c, gi, dens_ptr


_grad_outputs = None
_camera_out = None


def set_gradient_outputs(tensors, camera_out=None) -> None:
    """Where `backward` writes the six gradients: None (default) = freshly allocated tensors, as the reference does
    (rasterization_api.cu:127-134); or six preallocated contiguous CUDA float tensors of the parameters' shapes,
    e.g. views of a symmetric-memory buffer that peer GPUs read (fgs_b200.parallel.PeerAdam).  The returned
    gradients ARE these tensors (autograd then stores them as .grad without a copy).  The sixth entry (sh_rest) may be
    None: that gradient is then not computed at all and `None` is returned for it (compact view-parallel exchange).
    `camera_out`: a 3-float CUDA tensor that receives the view's camera position at every backward."""
    global _grad_outputs, _camera_out
    if tensors is not None:
        tensors = list(tensors)
        if len(tensors) != 6:
            raise ValueError('set_gradient_outputs: six tensors expected (means, scales, rotations, opacities, sh0, sh_rest)')
        for i, t in enumerate(tensors):
            if t is None and i == 5:
                continue
            _check_param(t, 'gradient output')
    _grad_outputs = tensors
    _camera_out = camera_out if tensors is not None else None


def backward(densification_info, grad_image, image, means, scales, rotations, opacities, sh_coefficients_rest,
             primitive_buffers, tile_buffers, instance_buffers, bucket_buffers, w2c, cam_position, bg_color,
             active_sh_bases, width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
             proper_antialiasing, n_instances, n_buckets, instance_primitive_indices_selector):
    """rasterization_api.h:30-59.  Returns the six gradient tensors."""
    dev, st, state, w2c_c, pos_c, bg_c, gi, dens_ptr = _backward_common(
        image, primitive_buffers, w2c, cam_position, bg_color, grad_image, densification_info, sh_coefficients_rest,
        active_sh_bases, width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
        proper_antialiasing, n_instances)
    with _capi.device_guard(dev):
        n = means.shape[0]
        if _grad_outputs is None:
            grads = [None if p is None else torch.empty_like(p) for p in (means, scales, rotations, opacities)]
            g_sh0 = torch.empty((n, 1, 3), dtype=torch.float32, device=dev)
            g_rest = torch.empty_like(sh_coefficients_rest)
        else:
            # fresh tensor objects over the same memory: autograd adopts a gradient as .grad without a copy only if
            # nobody else holds a reference to the object it was handed
            outs = [None if t is None else t.view(t.shape) for t in _grad_outputs]
            grads, g_sh0, g_rest = outs[:4], outs[4], outs[5]
            if _camera_out is not None:
                _camera_out.copy_(pos_c.reshape(-1)[:3])
            if rotations is None:
                raise RuntimeError('set_gradient_outputs is not available in covariance-input mode')
            shapes = [tuple(p.shape) for p in (means, scales, rotations, opacities)] + [(n, 1, 3), tuple(sh_coefficients_rest.shape)]
            for t, shape in zip(_grad_outputs, shapes):
                if t is None:
                    continue
                if tuple(t.shape) != shape or t.device != dev:
                    raise RuntimeError(f'gradient output has shape {tuple(t.shape)} on {t.device}, expected {shape} on {dev}')
        scratch = torch.empty(_lib.fgs_grad2d_scratch_bytes(n), dtype=torch.uint8, device=dev)
        rc = _lib.fgs_backward(_ptr(gi), _ptr(image), _ptr(means), _ptr(scales), _ptr(rotations), _ptr(opacities),
                               _ptr(sh_coefficients_rest), _ptr(w2c_c), _ptr(pos_c), _ptr(bg_c),
                               _ptr(primitive_buffers), _ptr(tile_buffers), _ptr(instance_buffers), _ptr(grads[0]),
                               _ptr(grads[1]), _ptr(grads[2]), _ptr(grads[3]), _ptr(g_sh0), _ptr(g_rest),
                               _ptr(scratch), dens_ptr, n, C.byref(st), C.byref(state), _stream())
        _capi.check(rc, 'backward')
    return grads[0], grads[1], grads[2], grads[3], g_sh0, g_rest


_deferred_keepalive: dict = {}  # device -> tensors the pending deferred update of that device still reads


def join_deferred(device=None) -> None:
    """Make the current stream wait for a pending deferred sh_rest update (stream-ordered, no host wait).  Call it before
    touching sh_coefficients_rest or its Adam moments with anything but this package's operators."""
    with _capi.device_guard(device if device is not None else torch.cuda.current_device()):
        _capi.check(_lib.fgs_join_deferred(_stream()), 'join_deferred')


def backward_adam(densification_info, grad_image, image, params, exp_avgs, exp_avg_sqs, steps, lrs, beta1, beta2, eps,
                  primitive_buffers, tile_buffers, instance_buffers, w2c, cam_position, bg_color, active_sh_bases,
                  width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane, proper_antialiasing,
                  n_instances, defer_rest=False) -> None:
    """FasterGSFused semantics: backward + FusedAdam.step in one pass (include/fgs_b200.h).
    `params`, `exp_avgs`, `exp_avg_sqs`: six tensors each in the order means, scales, rotations,
    opacities, sh_coefficients_0, sh_coefficients_rest; updated in place.
    `defer_rest`: the sh_rest part of the update runs on the library's side stream (include/fgs_b200.h "deferred SH
    schedule"); the next `forward(..., defer_color=True)` overlaps it with its geometry pass and sorts."""
    dev, st, state, w2c_c, pos_c, bg_c, gi, dens_ptr = _backward_common(
        image, primitive_buffers, w2c, cam_position, bg_color, grad_image, densification_info, params[5],
        active_sh_bases, width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
        proper_antialiasing, n_instances)
    for group in (params, exp_avgs, exp_avg_sqs):
        for t in group:
            _check_param(t, 'param/state')
    with _capi.device_guard(dev):
        n = params[0].shape[0]
        scratch = torch.empty(_lib.fgs_grad2d_scratch_bytes(n), dtype=torch.uint8, device=dev)
        arr = lambda ts: (C.c_void_p * 6)(*[_ptr(t) for t in ts])
        args = (_ptr(gi), _ptr(image), arr(params), arr(exp_avgs), arr(exp_avg_sqs),
                (C.c_int32 * 6)(*[int(s) for s in steps]),
                (C.c_double * 6)(*[float(l) for l in lrs]), float(beta1), float(beta2), float(eps),
                _ptr(w2c_c), _ptr(pos_c), _ptr(bg_c), _ptr(primitive_buffers), _ptr(tile_buffers),
                _ptr(instance_buffers), _ptr(scratch), dens_ptr, n, C.byref(st), C.byref(state), _stream())
        if defer_rest:
            rc = _lib.fgs_backward_adam_ex(*args, 1)
            # the side-stream kernel reads the 2D gradient scratch and the primitive buffers after this call has returned:
            # they stay referenced until the NEXT deferred backward on this device, by which time the forward in between
            # has made the caller's stream wait for the update (so the allocator may hand the memory out again)
            _deferred_keepalive[dev] = (scratch, primitive_buffers, params[5], exp_avgs[5], exp_avg_sqs[5])
        else:
            rc = _lib.fgs_backward_adam(*args)
        _capi.check(rc, 'backward_adam')


def inference(means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest, w2c, cam_position,
              bg_color, active_sh_bases, width, height, focal_x, focal_y, center_x, center_y, near_plane, far_plane,
              proper_antialiasing, to_chw, clamp_output):
    """rasterization_api.h:61-83."""
    for t, name in ((means, 'means'), (scales, 'scales'), (rotations, 'rotations'), (opacities, 'opacities'),
                    (sh_coefficients_0, 'sh_coefficients_0'), (sh_coefficients_rest, 'sh_coefficients_rest')):
        _check_param(t, name)
    dev = means.device
    with _capi.device_guard(dev):
        n = means.shape[0]
        total_rest = sh_coefficients_rest.shape[1] if sh_coefficients_rest.dim() > 1 else 0
        shape = (3, int(height), int(width)) if to_chw else (int(height), int(width), 3)
        image = torch.empty(shape, dtype=torch.float32, device=dev)
        w2c_c, pos_c, bg_c = _aux(w2c, dev), _aux(cam_position, dev), _aux(bg_color, dev)
        st = _settings(active_sh_bases, total_rest, width, height, focal_x, focal_y, center_x, center_y, near_plane,
                       far_plane, proper_antialiasing)
        with _Scratch(dev) as ctx:
            rc = _lib.fgs_inference(_resize_cb, ctx.user(_PRIMITIVE), ctx.user(_TILE), ctx.user(_INSTANCE), _ptr(means),
                                    _ptr(scales), _ptr(rotations), _ptr(opacities), _ptr(sh_coefficients_0),
                                    _ptr(sh_coefficients_rest), _ptr(w2c_c), _ptr(pos_c), _ptr(bg_c), _ptr(image), n,
                                    C.byref(st), int(bool(to_chw)), int(bool(clamp_output)), _stream())
            _capi.check(rc, 'inference')
        # the scratch tensors die here; the caching allocator keeps their memory stream-ordered behind the kernels
    return image


def adam_step(param_grad, param, exp_avg, exp_avg_sq, step_count, learning_rate, beta1, beta2, epsilon) -> None:
    """adam/include/adam.h:7-16."""
    for t, name in ((param_grad, 'param_grad'), (param, 'param'), (exp_avg, 'exp_avg'), (exp_avg_sq, 'exp_avg_sq')):
        _check_param(t, name)
    with _capi.device_guard(param.device):
        rc = _lib.fgs_adam_step(_ptr(param_grad), _ptr(param), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(),
                                int(step_count), float(learning_rate), float(beta1), float(beta2), float(epsilon),
                                _stream())
        _capi.check(rc, 'adam_step')


def update_3d_filter(positions, w2c, filter_3d, visibility_mask, width, height, focal_x, focal_y, center_x, center_y,
                     near_plane, clipping_tolerance, distance2filter) -> None:
    """filter3d/include/filter3d.h:7-20; filter_3d and visibility_mask (bool) are updated in place."""
    _check_param(positions, 'positions')
    _check_param(filter_3d, 'filter_3d')
    if not (_capi.on_device(visibility_mask) and visibility_mask.dtype == torch.bool and visibility_mask.is_contiguous()):
        raise RuntimeError("'visibility_mask' must be a contiguous CUDA bool tensor.")
    dev = positions.device
    with _capi.device_guard(dev):
        w2c_c = _aux(w2c, dev)
        rc = _lib.fgs_update_3d_filter(_ptr(positions), _ptr(w2c_c), _ptr(filter_3d), _ptr(visibility_mask),
                                       positions.shape[0], int(width), int(height), float(focal_x), float(focal_y),
                                       float(center_x), float(center_y), float(near_plane), float(clipping_tolerance),
                                       float(distance2filter), _stream())
        _capi.check(rc, 'update_3d_filter')


def relocation_adjustment(old_opacities, old_scales, n_samples_per_primitive):
    """densification_api.h:8-13.  Returns (new_opacities (n,1), new_scales (n,3))."""
    dev = old_opacities.device
    with _capi.device_guard(dev):
        n = old_opacities.shape[0]
        o = old_opacities.to(torch.float32).contiguous()
        sc = old_scales.to(torch.float32).contiguous()
        ns = n_samples_per_primitive.to(device=dev, dtype=torch.int64).contiguous()
        new_o = torch.empty((n, 1), dtype=torch.float32, device=dev)
        new_s = torch.empty((n, 3), dtype=torch.float32, device=dev)
        rc = _lib.fgs_relocation_adjustment(_ptr(o), _ptr(sc), _ptr(ns), _ptr(new_o), _ptr(new_s), n, _stream())
        _capi.check(rc, 'relocation_adjustment')
    return new_o, new_s


def add_noise(raw_scales, raw_rotations, raw_opacities, random_samples, means, current_lr) -> None:
    """densification_api.h:15-21; `means` is updated in place."""
    for t, name in ((raw_scales, 'raw_scales'), (raw_rotations, 'raw_rotations'), (raw_opacities, 'raw_opacities'),
                    (random_samples, 'random_samples'), (means, 'means')):
        _check_param(t, name)
    with _capi.device_guard(means.device):
        rc = _lib.fgs_add_noise(_ptr(raw_scales), _ptr(raw_rotations), _ptr(raw_opacities), _ptr(random_samples),
                                _ptr(means), raw_scales.shape[0], float(current_lr), _stream())
        _capi.check(rc, 'add_noise')


def pruning_scores(scores, means, scales, rotations, opacities, sh_coefficients_0, sh_coefficients_rest, w2c,
                   cam_position, bg_color, active_sh_bases, width, height, focal_x, focal_y, center_x, center_y,
                   near_plane, far_plane, proper_antialiasing) -> None:
    """rasterization_api.h:85-106: accumulates this view's Speedy-Splat pruning scores into `scores` (n,)."""
    for t, name in ((means, 'means'), (scales, 'scales'), (rotations, 'rotations'), (opacities, 'opacities'),
                    (sh_coefficients_0, 'sh_coefficients_0'), (sh_coefficients_rest, 'sh_coefficients_rest'),
                    (scores, 'scores')):
        _check_param(t, name)
    dev = means.device
    with _capi.device_guard(dev):
        n = means.shape[0]
        total_rest = sh_coefficients_rest.shape[1] if sh_coefficients_rest.dim() > 1 else 0
        image = torch.empty((3, int(height), int(width)), dtype=torch.float32, device=dev)
        w2c_c, pos_c, bg_c = _aux(w2c, dev), _aux(cam_position, dev), _aux(bg_color, dev)
        st = _settings(active_sh_bases, total_rest, width, height, focal_x, focal_y, center_x, center_y, near_plane,
                       far_plane, proper_antialiasing)
        with _Scratch(dev) as ctx:
            rc = _lib.fgs_pruning_scores(_resize_cb, ctx.user(_PRIMITIVE), ctx.user(_TILE), ctx.user(_INSTANCE),
                                         _ptr(means), _ptr(scales), _ptr(rotations), _ptr(opacities),
                                         _ptr(sh_coefficients_0), _ptr(sh_coefficients_rest), _ptr(w2c_c), _ptr(pos_c),
                                         _ptr(bg_c), _ptr(image), _ptr(scores), n, C.byref(st), _stream())
            _capi.check(rc, 'pruning_scores')
