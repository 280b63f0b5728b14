"""ctypes binding of ptt_b200/libptt_b200.so -- the C ABI declared in include/ptt_b200.h.

There is no fallback: if the library is missing (not built) `lib()` raises, and every op in
ptt_b200.ops raises on any non-zero return code.  Build with `python -m ptt_b200.build`.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libptt_b200.so")

_P = c_void_p           # device pointers travel as integers
_PP = ctypes.POINTER(c_void_p)
_IP = ctypes.POINTER(c_int)

# name -> (restype, argtypes); mirrors include/ptt_b200.h declaration by declaration
SIGNATURES = {
    "ptt_version": (c_char_p, []),
    "ptt_error_string": (c_char_p, [c_int]),
    "ptt_launch_count": (ctypes.c_ulonglong, []),
    "ptt_fault_status": (c_int, []),
    "ptt_fault_clear": (None, []),
    "ptt_furthest_point_sampling_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ptt_furthest_point_sampling": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "ptt_furthest_point_sampling_with_dist_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ptt_furthest_point_sampling_with_dist": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "ptt_gather_points": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "ptt_gather_points_grad": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "ptt_ball_query": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_int, _P, _P]),
    "ptt_ball_query_nested": (c_int, [_P, _P, c_int, c_int, c_int, _IP, ctypes.POINTER(c_float), _IP, _PP, _P]),
    "ptt_group_points": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "ptt_group_points_grad": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "ptt_three_nn": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "ptt_three_interpolate": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "ptt_three_interpolate_grad": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "ptt_cm_to_pm": (c_int, [_P, c_int, c_int, c_int, _P, c_int, _P]),
    "ptt_pm_to_cm": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "ptt_sa_params_floats": (c_size_t, [c_int, c_int, _IP]),
    "ptt_sa_pack_params": (c_int, [c_int, c_int, _IP, _PP, _PP, _PP, _P, _P]),
    "ptt_sa_mlp_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, _IP]),
    "ptt_sa_mlp_fwd": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                               _IP, _P, _P, c_int, _P, _P, c_size_t, _P]),
    "ptt_cosine_fusion_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, _IP]),
    "ptt_cosine_fusion_fwd": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _IP, _P, _P, c_int, _P,
                                      c_size_t, _P]),
    "ptt_knn": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "ptt_linear_params_floats": (c_size_t, [c_int, c_int]),
    "ptt_linear_pack": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "ptt_linear_pack_strided": (c_int, [_P, ctypes.c_longlong, ctypes.c_longlong, _P, c_int, c_int, _P, _P]),
    "ptt_linear_pack_batch": (c_int, [_P, c_int, _P]),
    "ptt_linear_fwd": (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int, _P, c_int, _P]),
    "ptt_transformer_params_floats": (c_size_t, [c_int, c_int]),
    "ptt_transformer_pack_params": (c_int, [c_int, c_int] + [_P] * 15 + [_P, _P]),
    "ptt_transformer_block_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "ptt_transformer_block_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                          c_size_t, _P]),
    "ptt_transformer_block_fwd_ex": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P,
                                             _P, _P, c_size_t, _P]),
    "ptt_pair_cosine": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "ptt_layer_norm_fwd": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_float, _P, c_int, _P, c_int, _P]),
    "ptt_token_softmax_gate": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_float, _P, c_int, _P, _P]),
    "ptt_transformer_std_params_floats": (c_size_t, [c_int, c_int]),
    "ptt_transformer_std_pack_params": (c_int, [c_int, c_int] + [_P] * 11 + [_P, _P]),
    "ptt_transformer_std_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ptt_transformer_std_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "ptt_linear_fwd_ex": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P, c_int, _P, c_int, _P]),
    "ptt_linear_fwd_stats": (c_int, [_P, c_int, ctypes.c_longlong, c_int, _P, _P, _P, c_int, c_int, _P, c_int, _P, _P]),
    "ptt_linear_wgrad": (c_int, [_P, c_int, _P, c_int, _P, _P, ctypes.c_longlong, c_int, c_int, _P, c_int, _P, _P]),
    "ptt_col_stats": (c_int, [_P, c_int, ctypes.c_longlong, c_int, _P, _P]),
    "ptt_bn_train_finalize": (c_int, [_P, ctypes.c_longlong, c_int, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P]),
    "ptt_sa_group_rows": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P, c_int, _P]),
    "ptt_sa_group_rows_grad": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P, c_int, _P, _P, _P]),
    "ptt_bn_relu_maxpool": (c_int, [_P, c_int, ctypes.c_longlong, c_int, c_int, _P, _P, _P, c_int, _P, _P]),
    "ptt_bn_relu_bwd": (c_int, [_P, c_int, _P, c_int, _P, c_int, ctypes.c_longlong, c_int, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P]),
    "ptt_transformer_block_workspace_layout": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "ptt_tr_softmax_bwd": (c_int, [_P, c_int, _P, _P, c_int, ctypes.c_longlong, c_int, c_int, c_float, _P, _P, _P]),
    "ptt_tr_pair_inputs": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, c_int, _P, _P, _P, _P]),
    "ptt_tr_mask_positive": (c_int, [_P, _P, ctypes.c_longlong, _P]),
    "ptt_tr_rows_linear": (c_int, [_P, c_int, ctypes.c_longlong, c_int, _P, _P, _P, _P]),
    "ptt_tr_pair_scatter": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, _P]),
    "ptt_mt19937_stream": (c_int, [ctypes.c_uint, c_int, _P]),
    "ptt_track_crop": (c_int, [c_int, c_int, _PP, _PP, _PP, _IP, _IP, ctypes.c_double, ctypes.c_double, c_int, _P, c_int, _P, _P]),
    "ptt_track_regularize": (c_int, [_P, _P, c_int, c_int, c_int, _P, c_int, _P, _P, _P]),
    "ptt_track_update": (c_int, [_P, c_int, _P, c_int, c_int, _P, c_int, _P, _P, c_int, _P, _P]),
}

# include/ptt_b200_tuning.h: process-wide test / tuning switches.  Bound for tests/ and tools/ only -- nothing under
# ptt_b200/ calls them (tests/test_abi.py checks that).
TUNING_SIGNATURES = {
    "ptt_debug_force_ffma": (None, [c_int]),
    "ptt_debug_set_cluster": (None, [c_int]),
    "ptt_debug_sa_timeline": (None, [_P]),
    "ptt_debug_tr_pass_timeline": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, _P, _P, c_int, _P, c_int, _P, _P]),
    "ptt_fps_variant": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, _P]),
}

_lib = None


class PttError(RuntimeError):
    """A ptt_* entry point returned a non-zero code (the reference's `_ext` raises RuntimeError too)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PttError("%s is missing: build it with `python -m ptt_b200.build` (there is no fallback path)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in list(SIGNATURES.items()) + list(TUNING_SIGNATURES.items()):
            fn = getattr(handle, name)   # AttributeError if the library does not export what the header declares
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


# return codes of include/ptt_b200.h
PTT_OK, PTT_ERR_INVALID_ARGUMENT, PTT_ERR_UNSUPPORTED, PTT_ERR_WORKSPACE, PTT_ERR_DEVICE_FAULT = 0, -1, -2, -3, -4


def check(code, what):
    if code != 0:
        raise PttError("%s failed: %s (code %d)" % (what, lib().ptt_error_string(code).decode(), code))


def fault_status():
    """0, or PTT_ERR_DEVICE_FAULT once a kernel gave up a bounded barrier wait (sticky until fault_clear())."""
    return int(lib().ptt_fault_status())


def fault_clear():
    lib().ptt_fault_clear()


def launch_count():
    """Kernels launched by libptt_b200.so in this process so far."""
    return int(lib().ptt_launch_count())


def version():
    return lib().ptt_version().decode()
