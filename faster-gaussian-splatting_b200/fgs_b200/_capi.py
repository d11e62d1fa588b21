"""ctypes binding of libfgs_b200.so (the C ABI declared in include/fgs_b200.h).

The library is the product: there is no Python/torch fallback.  If it cannot be loaded the import
fails loudly (the reference raises Framework.ExtensionError in the same situation,
/root/reference/FasterGSCudaBackend/__init__.py:19-20).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / 'lib' / 'libfgs_b200.so'

ResizeFn = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class Settings(C.Structure):
    """struct fgs_settings"""
    _fields_ = [
        ('active_sh_bases', C.c_int32),
        ('total_sh_bases_rest', C.c_int32),
        ('width', C.c_int32),
        ('height', C.c_int32),
        ('focal_x', C.c_float),
        ('focal_y', C.c_float),
        ('center_x', C.c_float),
        ('center_y', C.c_float),
        ('near_plane', C.c_float),
        ('far_plane', C.c_float),
        ('proper_antialiasing', C.c_int32),
    ]


class ForwardState(C.Structure):
    """struct fgs_forward_state"""
    _fields_ = [
        ('n_instances', C.c_int32),
        ('n_buckets', C.c_int32),
        ('selector', C.c_int32),
        ('n_visible', C.c_int32),
    ]


class RowTensor(C.Structure):
    """struct fgs_row_tensor"""
    _fields_ = [('src', C.c_void_p), ('appended', C.c_void_p), ('dst', C.c_void_p), ('row_floats', C.c_int32),
                ('role', C.c_int32)]


_P = C.c_void_p
_PtrArr6 = C.c_void_p * 6
_I32Arr6 = C.c_int32 * 6
_F64Arr6 = C.c_double * 6
_U64Arr6 = C.c_uint64 * 6

# name -> (restype, argtypes); must list every function declared in include/fgs_b200.h
SIGNATURES = {
    'fgs_abi_version': (C.c_int, []),
    'fgs_last_error': (C.c_char_p, []),
    'fgs_forward': (C.c_int, [ResizeFn, _P, _P, _P] + [_P] * 9 + [_P, C.c_int32, C.POINTER(Settings), _P,
                                                                    C.POINTER(ForwardState)]),
    'fgs_grad2d_scratch_bytes': (C.c_size_t, [C.c_int32]),
    'fgs_backward': (C.c_int, [_P] * 10 + [_P] * 3 + [_P] * 6 + [_P, _P, C.c_int32, C.POINTER(Settings),
                                                                    C.POINTER(ForwardState), _P]),
    'fgs_backward_adam': (C.c_int, [_P, _P, _PtrArr6, _PtrArr6, _PtrArr6, _I32Arr6, _F64Arr6, C.c_double,
                                     C.c_double, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32,
                                     C.POINTER(Settings), C.POINTER(ForwardState), _P]),
    'fgs_forward_ex': (C.c_int, [ResizeFn, _P, _P, _P] + [_P] * 9 + [_P, C.c_int32, C.POINTER(Settings), _P,
                                                                       C.POINTER(ForwardState), C.c_uint32, _P]),
    'fgs_backward_adam_ex': (C.c_int, [_P, _P, _PtrArr6, _PtrArr6, _PtrArr6, _I32Arr6, _F64Arr6, C.c_double,
                                        C.c_double, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32,
                                        C.POINTER(Settings), C.POINTER(ForwardState), _P, C.c_uint32]),
    'fgs_join_deferred': (C.c_int, [_P]),
    'fgs_debug_set_deferred_ctas_per_sm': (None, [C.c_int32]),
    'fgs_inference': (C.c_int, [ResizeFn, _P, _P, _P] + [_P] * 9 + [_P, C.c_int32, C.POINTER(Settings), C.c_int32,
                                                                      C.c_int32, _P]),
    'fgs_adam_step': (C.c_int, [_P, _P, _P, _P, C.c_size_t, C.c_int32, C.c_double, C.c_double, C.c_double,
                                 C.c_double, _P]),
    'fgs_pruning_scores': (C.c_int, [ResizeFn, _P, _P, _P] + [_P] * 9 + [_P, _P, C.c_int32, C.POINTER(Settings), _P]),
    'fgs_update_3d_filter': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32] + [C.c_float] * 7 + [_P]),
    'fgs_relocation_adjustment': (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, _P]),
    'fgs_add_noise': (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_float, _P]),
    'fgs_loss_scratch_bytes': (C.c_size_t, [C.c_int32, C.c_int32]),
    'fgs_loss_forward': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P, _P]),
    'fgs_loss_backward': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P, _P, _P]),
    'fgs_loss_forward_backward': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P, _P,
                                             _P]),
    'fgs_gather_rows': (C.c_int, [C.POINTER(RowTensor), C.c_int32, _P, C.c_int64, C.c_int64, _P, _P, _P, _P]),
    'fgs_zero_rows': (C.c_int, [_P, _P, C.c_int32, _P, C.c_int64, _P]),
    'fgs_densify_scratch_bytes': (C.c_size_t, [C.c_int32]),
    'fgs_plan_count': (C.c_int, [_P, C.c_int32, _P, C.POINTER(C.c_uint32), _P]),
    'fgs_plan_emit': (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    'fgs_adc_flags': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_float, _P, _P]),
    'fgs_morton_keys': (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    'fgs_gather_u32': (C.c_int, [_P, _P, _P, _P, C.c_int32, _P]),
    'fgs_slice4d_forward': (C.c_int, [_P, _P, _P, _P, _P, C.c_float, C.c_int32, _P, _P, _P, _P]),
    'fgs_slice4d_backward': (C.c_int, [_P, _P, _P, _P, _P, C.c_float, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'fgs_peer_adam_step': (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _U64Arr6, _U64Arr6, _I32Arr6, _F64Arr6,
                                      C.c_double, C.c_double, C.c_double, _P]),
    'fgs_peer_adam_step_multicast': (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _U64Arr6, _U64Arr6, _I32Arr6,
                                                _F64Arr6, C.c_double, C.c_double, C.c_double, _P]),
    'fgs_peer_adam_rows_per_rank': (C.c_uint64, [C.c_uint64, C.c_int32]),
    'fgs_peer_adam_rows_step': (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _U64Arr6, C.c_uint64, C.c_int32, C.c_int32,
                                           C.c_uint64, C.c_int32, _I32Arr6, _F64Arr6, C.c_double, C.c_double, C.c_double, _P, _P, _P,
                                           C.c_int32, _P]),
    'fgs_peer_adam_small_step': (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _U64Arr6, C.c_uint64, C.c_int32, _I32Arr6,
                                            _F64Arr6, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P]),
    'fgs_peer_broadcast_multicast': (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_int32, _P]),
    'fgs_peer_adam_shard_floats': (C.c_size_t, [_U64Arr6, C.c_int32]),
    'fgs_peer_adam_slice': (C.c_int, [_U64Arr6, C.c_int32, C.c_int32, _U64Arr6, _U64Arr6, _U64Arr6]),
    'fgs_peer_last_error': (C.c_char_p, []),
    'fgs_debug_set_overlap_adam': (None, [C.c_int32]),
    'fgs_debug_get_overlap_adam': (C.c_int32, []),
    'fgs_debug_set_capacity_hint': (None, [C.c_int32]),
    'fgs_debug_get_capacity_hint': (C.c_int32, []),
    'fgs_debug_get_forward_reruns': (C.c_int32, []),
    'fgs_debug_blend_stats': (C.c_int, [C.POINTER(C.c_uint64), C.c_int32]),
    'fgs_profile_enable': (None, [C.c_int32]),
    'fgs_profile_read': (C.c_int, [C.c_char_p, C.c_size_t]),
    'fgs_sort_pairs_u32': (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_size_t,
                                      C.POINTER(C.c_int32), _P]),
    'fgs_sort_pairs_u16': (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_size_t,
                                      C.POINTER(C.c_int32), _P]),
    'fgs_sort_scratch_bytes': (C.c_size_t, [C.c_int32, C.c_int32]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the library (building it first if only the sources are present and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get('FGS_B200_LIB')  # experiment variants built with build.py --out
    if override:
        lib_path = Path(override)
        if not lib_path.exists():
            raise ImportError(f'fgs_b200: FGS_B200_LIB={override} does not exist')
    elif not LIB_PATH.exists() or os.environ.get('FGS_B200_REBUILD'):
        import importlib.util
        spec = importlib.util.spec_from_file_location('fgs_b200_build', _PKG.parent / 'build.py')
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build(force=bool(os.environ.get('FGS_B200_REBUILD')))
    if not override:
        lib_path = LIB_PATH
    if not lib_path.exists():
        raise ImportError(f'fgs_b200: {lib_path} is missing and could not be built; there is no fallback path')
    lib = C.CDLL(str(lib_path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so and the header disagree
        fn.restype = res
        fn.argtypes = args
    if lib.fgs_abi_version() != 1:
        raise ImportError('fgs_b200: ABI version mismatch between libfgs_b200.so and the Python binding')
    _lib = lib
    return lib


# The three places where the Python layer touches the CUDA runtime through torch.  (tests/test_emulated_python_layer.py
# replaces them to run this layer's argument marshalling on host tensors against the emulated build of the library.)
def on_device(t) -> bool:
    return bool(t.is_cuda)


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def device_guard(device):
    import torch
    return torch.cuda.device(device)


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().fgs_last_error().decode(errors='replace')
        raise RuntimeError(f'fgs_b200 {what} failed (status {status}): {msg}')


def profile_enable(on: bool) -> None:
    load().fgs_profile_enable(1 if on else 0)


def profile_read() -> list[tuple[str, float]]:
    """[(kernel name, milliseconds)] in launch order since the last read."""
    _last_starts.clear()
    buf = C.create_string_buffer(1 << 16)
    check(load().fgs_profile_read(buf, len(buf)), 'profile_read')
    out = []
    for line in buf.value.decode().splitlines():
        name, _, rest = line.partition('=')
        ms, _, t0 = rest.partition('@')
        out.append((name, float(ms)))
        _last_starts.append((name, float(t0 or 0.0), float(ms)))
    return out


_last_starts: list = []  # (name, start offset in ms since the first profiled launch, duration) of the last reads
