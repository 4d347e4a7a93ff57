"""ctypes binding of the C ABI declared in include/iif_b200.h (libiif_b200.so, built in-tree).

There is NO fallback: if the shared library is missing and cannot be built, or a call returns a
non-zero code, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libiif_b200.so")

OK, EINVAL, EALIGN, EUNSUPPORTED, EWORKSPACE, EDRIVER = 0, -1, -2, -3, -4, -5
VARIANT_IDS = {"raw": 0, "smooth": 1, "rel": 2, "prob": 2, "normit": 3, "gombit": 4, "base2": 5, "base10": 6}
DTYPE_F32, DTYPE_BF16 = 0, 1
HEAD_NO_FUSED_LOSS = 1
HEAD_NO_PERSISTENT = 2
HEAD_LOW_REGS = 4
NORM_NORMED, NORM_COS, NORM_UNIT = 0, 1, 2

_p, _i64, _i32, _f32, _f64, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double, C.c_size_t


class HeadArgs(C.Structure):
    """struct iif_head_args (include/iif_b200.h)."""
    _fields_ = [
        ("x", _p), ("ldx", _i64), ("w", _p), ("ldw", _i64), ("bias", _p), ("iif", _p), ("label", _p),
        ("class_weight", _p), ("sample_weight", _p), ("ignore_index", _i64), ("scale", _f32),
        ("B", _i64), ("D", _i64), ("C", _i64),
        ("z", _p), ("ldz", _i64), ("loss_i", _p), ("loss_sum", _p), ("dz_bf16", _p), ("lddz", _i64),
        ("dx", _p), ("dx_dtype", _i32), ("lddx", _i64), ("dw", _p), ("lddw", _i64), ("db", _p),
        ("argmax", _p), ("rank", _p), ("acc_counts", _p), ("scratch", _p), ("ws", _p), ("ws_bytes", _sz), ("flags", _i32),
    ]


# name -> (restype, argtypes); mirrors include/iif_b200.h one to one
SIGNATURES = {
    "iif_abi_version": (_i32, []),
    "iif_error_string": (C.c_char_p, [_i32]),
    "iif_launch_count": (C.c_uint64, []),
    "iif_hist_labels_i64": (_i32, [_p, _i64, _p, _i64, _p]),
    "iif_hist_images_dedup_i64": (_i32, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "iif_hist_images_dedup_ws_bytes": (_sz, [_i64, _i64]),
    "iif_weights_from_counts": (_i32, [_p, _i64, _i64, _i32, _f64, _p, _p, _p]),
    "iif_softmax_ce_fwd_bwd": (_i32, [_p, _i64, _p, _p, _p, _p, _i64, _f32, _i64, _i64, _p, _p, _p, _i64, _p, _i64,
                                      _p, _p, _p, _p, _p, _p]),
    "iif_loss_scratch_bytes": (_sz, [_i64]),
    "iif_softmax_ce_mixup_fwd_bwd": (_i32, [_p, _i64, _p, _p, _p, _f32, _p, _p, _i64, _f32, _i64, _i64, _p, _p, _p, _i64,
                                            _p, _i64, _p, _p, _p, _p, _p]),
    "iif_scaled_activation": (_i32, [_p, _i64, _p, _i32, _i64, _i64, _p, _i64, _p, _p, _p, _p]),
    "iif_sigmoid_bce_fwd_bwd": (_i32, [_p, _i64, _p, _p, _p, _p, _i64, _f32, _i64, _i64, _p, _i64, _p, _p, _p, _i64,
                                       _p, _i64, _p, _p]),
    "iif_sigmoid_focal_fwd_bwd": (_i32, [_p, _i64, _p, _f32, _f32, _p, _p, _i64, _f32, _i64, _i64, _p, _i64, _p, _p, _p, _i64,
                                         _p, _i64, _p, _p]),
    "iif_sigmoid_bce_dense_fwd_bwd": (_i32, [_p, _i64, _p, _i64, _p, _p, _i64, _f32, _i64, _i64, _p, _i64, _p, _p, _p, _i64,
                                             _p, _p]),
    "iif_class_accumulate": (_i32, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "iif_class_feature_stats": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _f32, _p, _p, _i64, _p, _p, _p]),
    "iif_class_feature_stats_ws_bytes": (_sz, [_i64]),
    "iif_shot_accuracy": (_i32, [_p, _p, _i64, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p]),
    "iif_row_scale_from_norm": (_i32, [_p, _i64, _i64, _i64, _p, _i32, _f32, _f32, _f32, _p, _p, _p, _p]),
    "iif_row_dot": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _p, _p]),
    "iif_rows_axpby": (_i32, [_p, _i64, _p, _p, _i64, _p, _p, _i64, _i64, _p, _i64, _p]),
    "iif_scale_rows": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _p, _i32, _i64, _p]),
    "iif_scale_inplace": (_i32, [_p, _i32, _i64, _p, _p]),
    "iif_colsum": (_i32, [_p, _i32, _i64, _p, _i64, _i64, _p, _p]),
    "iif_linear_fwd_bf16": (_i32, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _p, _i64, _i64, _i64, _i64, _p, _sz, _p]),
    "iif_linear_fwd_f32": (_i32, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _p, _i64, _i64, _i64, _i64, _p]),
    "iif_split3_bf16": (_i32, [_p, _i64, _i64, _i64, _i32, _i32, _p, _i64, _p]),
    "iif_linear_bwd_dx_bf16": (_i32, [_p, _i64, _p, _i64, _p, _p, _i32, _i64, _i64, _i64, _i64, _p, _sz, _p]),
    "iif_linear_bwd_dx_f32": (_i32, [_p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _p]),
    "iif_linear_bwd_dw_bf16": (_i32, [_p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _p, _sz, _p]),
    "iif_linear_bwd_dw_f32": (_i32, [_p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _p]),
    "iif_linear_bwd_bf16": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _p, _i32, _i64, _p, _i64, _p, _i64, _i64, _i64, _p,
                                   _sz, _p]),
    "iif_debug_timing": (None, [_p]),
    "iif_debug_timing_fused": (None, [_p]),
    "iif_debug_fused_plan": (_i32, [_i64, _i64, _i64, _i32, _i32, _p]),
    "iif_debug_timing_allreduce": (None, [_p]),
    "iif_debug_capacity": (_i32, [_p]),
    "iif_allreduce_mean_f32": (_i32, [_p, _p, _p, _i32, _i32, _i64, _i64, _i32, _i32, _i32, _p]),
    "iif_allreduce_flag_bytes": (_sz, []),
    "iif_allreduce_set_timeout_ms": (_i32, [_i64]),
    "iif_pipeline_create": (_i32, [C.POINTER(_p), C.POINTER(HeadArgs), _i32]),
    "iif_pipeline_submit": (_i32, [_p, _i32, _p, _p, _p]),
    "iif_pipeline_submit_device": (_i32, [_p, _i32]),
    "iif_pipeline_join": (_i32, [_p, _p]),
    "iif_pipeline_set_allreduce": (_i32, [_p, _p, _p, _p, _i32, _i32, C.POINTER(_i64), _i64, _i32, _i32, _i32]),
    "iif_pipeline_get_streams": (_i32, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p)]),
    "iif_pipeline_enable_staged": (_i32, [_p]),
    "iif_pipeline_staging": (_i32, [_p, _i32, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p)]),
    "iif_pipeline_submit_staged": (_i32, [_p, _i32]),
    "iif_pipeline_submit_staged_ring": (_i32, [_p]),
    "iif_pipeline_wait": (_i32, [_p, _i32]),
    "iif_pipeline_stream_wait_step": (_i32, [_p, _i32, _p]),
    "iif_pipeline_hold_slot": (_i32, [_p, _i32, _p]),
    "iif_pipeline_sync": (_i32, [_p]),
    "iif_pipeline_destroy": (None, [_p]),
    "iif_gemm_ws_bytes": (_sz, [_i64, _i64, _i64]),
    "iif_gemm_reserve_slots": (_i32, [_i32]),
    "iif_gemm_assume_exclusive": (_i32, [_i32]),
    "iif_head_fwd_bwd_bf16": (_i32, [C.POINTER(HeadArgs), _p]),
    "iif_loss_linear_bwd_bf16": (_i32, [C.POINTER(HeadArgs), _p]),
    "iif_head_launches": (_i32, [C.POINTER(HeadArgs)]),
}

_lib = None


def load():
    """Load (building first if the .so is absent and nvcc exists). Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.iif_abi_version() != 1:
        raise RuntimeError("libiif_b200.so ABI version mismatch; rebuild with `python -m iif_b200.build --force`")
    _lib = lib
    return lib


def error_string(code: int) -> str:
    return load().iif_error_string(int(code)).decode()


def check(code: int, what: str = "") -> None:
    if code != 0:
        raise RuntimeError(f"iif_b200 {what} failed: [{code}] {error_string(code)}")
