"""ctypes binding of libracformer_ops.so (declared in include/racformer_ops.h).

There is no CPU fallback: if the library is missing, loading raises. Nothing here imports `oracle/`.
"""
import ctypes
import os
import threading

from . import build as _build

_c_float_p = ctypes.c_void_p  # device pointers are passed as integers
_i = ctypes.c_int

# name -> (restype, argtypes); must match include/racformer_ops.h
SIGNATURES = {
    "racf_version": (_i, []),
    "racf_status_string": (ctypes.c_char_p, [_i]),
    "racf_msmv_forward": (_i, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_i), _i, _c_float_p, _c_float_p,
                               _i, _i, _i, _i, _i, _c_float_p, ctypes.c_void_p]),
    "racf_msmv_forward_variant": (_i, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_i), _i, _c_float_p, _c_float_p,
                                       _i, _i, _i, _i, _i, _i, _c_float_p, ctypes.c_void_p]),
    "racf_msmv_forward_grouped": (_i, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_i), _i, _c_float_p, _c_float_p,
                                       _i, _i, _i, _i, _i, _i, _i, _c_float_p, ctypes.c_void_p]),
    "racf_msmv_backward": (_i, [_c_float_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_i), _i, _c_float_p,
                                _c_float_p, _i, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_void_p), _c_float_p,
                                _c_float_p, _i, ctypes.c_void_p]),
    "racf_msmv_backward_grouped": (_i, [_c_float_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_i), _i, _c_float_p,
                                        _c_float_p, _i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_void_p), _c_float_p,
                                        _c_float_p, _i, ctypes.c_void_p]),
    "racf_msmv_tap_masks": (_i, [ctypes.POINTER(_i), _i, _c_float_p, _i, _i, _i, _i, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p]),
    "racf_msda_forward": (_i, [_c_float_p, ctypes.c_void_p, ctypes.c_void_p, _c_float_p, _c_float_p,
                               _i, _i, _i, _i, _i, _i, _i, _i, _c_float_p, ctypes.c_void_p]),
    "racf_msda_forward_pair": (_i, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                    ctypes.c_void_p, ctypes.c_void_p] + [_i] * 8 + [ctypes.c_void_p]),
    "racf_msda_backward": (_i, [_c_float_p, ctypes.c_void_p, ctypes.c_void_p, _c_float_p, _c_float_p, _c_float_p,
                                _i, _i, _i, _i, _i, _i, _i, _i, _c_float_p, _c_float_p, _c_float_p,
                                ctypes.c_void_p]),
    "racf_msmv_points_forward": (_i, [_c_float_p] * 7 + [ctypes.POINTER(ctypes.c_double)] + [ctypes.c_float] * 4
                                 + [_i] * 8 + [_c_float_p, _c_float_p, ctypes.c_void_p]),
    "racf_bev_points_forward": (_i, [_c_float_p] * 6 + [ctypes.POINTER(ctypes.c_double), ctypes.c_float] + [_i] * 6
                                + [_c_float_p, _c_float_p, ctypes.c_void_p]),
    "racf_to_sampling_layout": (_i, [_c_float_p, _c_float_p] + [_i] * 7 + [ctypes.c_void_p]),
    "racf_from_sampling_layout": (_i, [_c_float_p, _c_float_p] + [_i] * 7 + [ctypes.c_void_p]),
    "racf_to_sampling_layout_f16": (_i, [ctypes.c_void_p, _c_float_p] + [_i] * 8 + [ctypes.c_void_p]),
    "racf_bev_pool_v2_forward": (_i, [_c_float_p] * 7 + [_i, _i, _c_float_p, ctypes.c_void_p]),
    "racf_bev_pool_v2_backward": (_i, [_c_float_p] * 8 + [_i, _i, _c_float_p, _c_float_p, ctypes.c_void_p]),
    "racf_adaptive_mixing_forward": (_i, [_c_float_p, _c_float_p, _i, _i, _i, _i, ctypes.c_float, _c_float_p, ctypes.c_void_p]),
    "racf_adaptive_mixing_forward_split": (_i, [_c_float_p, _c_float_p, _i, _i, _i, _i, ctypes.c_float, ctypes.c_void_p,
                                                _i, ctypes.c_void_p]),
    "racf_adaptive_mixing_tc_forward": (_i, [_c_float_p, _c_float_p, _i, _i, _i, _i, ctypes.c_float, _c_float_p,
                                             ctypes.c_void_p, _i, ctypes.c_void_p]),
    "racf_adaptive_mixing_tc_forward_variant": (_i, [_c_float_p, _c_float_p, _i, _i, _i, _i, ctypes.c_float, _c_float_p,
                                                     ctypes.c_void_p, _i, _i, ctypes.c_void_p]),
    "racf_adaptive_mixing_backward": (_i, [_c_float_p, _c_float_p, _c_float_p, _i, _i, _i, _i, ctypes.c_float, _c_float_p,
                                           _c_float_p, ctypes.c_void_p]),
    "racf_adaptive_mixing_backward_variant": (_i, [_c_float_p, _c_float_p, _c_float_p, _i, _i, _i, _i, ctypes.c_float, _c_float_p,
                                                   _c_float_p, _i, ctypes.c_void_p]),
    "racf_linear_tiled_bytes": (ctypes.c_longlong, [ctypes.c_longlong, _i]),
    "racf_split_bf16x3_tiled": (_i, [_c_float_p, ctypes.c_longlong, _i, ctypes.c_void_p, ctypes.c_void_p]),
    "racf_split_bf16x3_tiled_add": (_i, [_c_float_p, ctypes.c_longlong, _i, _c_float_p, ctypes.c_longlong, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "racf_split_bf16x3": (_i, [_c_float_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p]),
    "racf_split_bf16x3_chw_to_hwc": (_i, [_c_float_p, _c_float_p, _i, _i, _i, _i, ctypes.c_void_p, ctypes.c_void_p]),
    "racf_linear_bf16x3_plan": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_longlong)]),
    "racf_linear_bf16x3_forward": (_i, [ctypes.c_void_p, ctypes.c_void_p, _c_float_p, _i, _i, _i, _i, _i, _i, _c_float_p,
                                        _c_float_p, ctypes.c_void_p]),
    "racf_linear_bf16x3_multi_forward": (_i, [ctypes.c_void_p, ctypes.c_void_p, _i, _i, _i, ctypes.POINTER(_i),
                                              ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), _i, _i,
                                              ctypes.c_void_p]),
    "racf_chw_to_hwc": (_i, [_c_float_p, _i, _i, _i, _c_float_p, _i, _c_float_p, _i, ctypes.c_void_p]),
    "racf_sasa_attention_forward": (_i, [_c_float_p, _c_float_p, _c_float_p, ctypes.POINTER(ctypes.c_double), _i, _i, _i, _i,
                                         _i, _c_float_p, ctypes.c_void_p]),
    "racf_sasa_attention_train_forward": (_i, [_c_float_p, _c_float_p, _c_float_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                               _i, _i, _i, _i, _i, ctypes.c_float, ctypes.c_uint, _c_float_p, _c_float_p,
                                               ctypes.c_void_p]),
    "racf_sasa_attention_train_backward": (_i, [_c_float_p, _c_float_p, _c_float_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                                _i, _i, _i, _i, _i, ctypes.c_float, ctypes.c_uint, _c_float_p, _c_float_p,
                                                _c_float_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_void_p]),
    "racf_refine_bbox_forward": (_i, [_c_float_p, _c_float_p, _c_float_p, _i, _i, _i, _i, ctypes.c_float, _c_float_p,
                                      _c_float_p, ctypes.c_void_p]),
    "racf_refine_bbox_backward": (_i, [_c_float_p] * 4 + [_i, _i, _i, _i, ctypes.c_float, _c_float_p, _c_float_p, ctypes.c_void_p]),
    "racf_msmv_points_backward": (_i, [_c_float_p] * 6 + [ctypes.POINTER(ctypes.c_double)] + [ctypes.c_float] * 4 + [_i] * 8
                                  + [_c_float_p] * 8 + [ctypes.c_void_p]),
    "racf_bev_points_backward": (_i, [_c_float_p] * 5 + [ctypes.POINTER(ctypes.c_double), ctypes.c_float] + [_i] * 6
                                 + [_c_float_p] * 7 + [ctypes.c_void_p]),
    "racf_convgru_gates_forward": (_i, [_c_float_p, _c_float_p, ctypes.c_longlong, _i, _c_float_p, ctypes.c_void_p]),
    "racf_upsample2x_bilinear_nhwc": (_i, [_c_float_p, _i, _i, _i, _i, _c_float_p, ctypes.c_void_p]),
    "racf_row_program_forward": (_i, [ctypes.c_void_p, _i, _i, _i, _i, _i, ctypes.c_void_p]),
    "racf_msda_tap_masks": (_i, [ctypes.c_void_p, _c_float_p, _i, _i, _i, _i, _i, ctypes.c_void_p, ctypes.c_void_p]),
}

_lock = threading.Lock()
_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Return the loaded library (ctypes.CDLL). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = lib_path()
            if not os.path.exists(path):
                raise ImportError(
                    f"{path} not found: build it with `python -m racformer_b200.build` "
                    "(there is no CPU fallback for the racformer_b200 ops)")
            lib = ctypes.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the symbol is missing
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def status_string(code):
    return load().racf_status_string(int(code)).decode()


class Unsupported(RuntimeError):
    """RACF_ERR_UNSUPPORTED: the fused variant does not exist for these shapes / alignments. Callers that have another
    GPU path for the same computation (e.g. the decoder's PyTorch operator chain) catch this and use it."""


CALLS = [0]   # C-ABI calls that went through check(): each launched at least one kernel (bench.py's launch count in training)


def check(code, what):
    """Turn a non-zero status into the RuntimeError the reference's ATen asserts would raise."""
    CALLS[0] += 1
    if code == -6:
        raise Unsupported(f"{what}: {status_string(code)} (status {code})")
    if code != 0:
        raise RuntimeError(f"{what}: {status_string(code)} (status {code})")
