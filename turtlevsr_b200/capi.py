"""ctypes binding of libturtle_b200.so (the C ABI declared in include/turtle_b200.h).

The library is mandatory: importing this module on a machine where the .so is missing tries to
build it with nvcc, and raises if that fails.  There is no CPU or torch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

MAX_SEG = 48
SAB_SLOTS = 48
FP32, TF32 = 0, 1
ACT_NONE, ACT_GELU = 0, 1
STORE_PLAIN, STORE_UNSHUFFLE2, STORE_SHUFFLE2 = 0, 1, 2
ENOTSUP = -3
METRICS_INFERENCE, METRICS_BASICSR, METRICS_FLOAT = 0, 1, 2

_fp = C.c_void_p
_i32, _i64 = C.c_int32, C.c_int64
_f32 = C.c_float


class GemmArgs(C.Structure):
    _fields_ = [
        ("mode", _i32), ("im2col", _i32), ("P", _i64), ("B", _i32), ("H", _i32), ("W", _i32), ("Cout", _i32),
        ("nseg", _i32), ("segw", _i32), ("A", _fp * MAX_SEG), ("lda", _i32 * MAX_SEG), ("Wt", _fp),
        ("bias", _fp), ("scale", _fp), ("act", _i32), ("res", _fp), ("ldres", _i32), ("out", _fp),
        ("ldo", _i32), ("store", _i32), ("round_out", _i32), ("a_dtype", _i32), ("out_dtype", _i32),
        ("ln_out", _fp), ("ld_ln", _i32), ("ln_w", _fp), ("ln_b", _fp),
        ("w_batches", _i32), ("reserved_", _i32), ("w_bstride", _i64), ("rows_per_batch", _i64),
    ]


_SIGS = {
    "turtle_abi_version": ([], C.c_int),
    "turtle_sizeof_gemm_args": ([], C.c_int),
    "turtle_build_info": ([], C.c_char_p),
    "turtle_pack_frame": ([_fp, _i64, _fp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_conv3x3_first": ([_fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_conv3x3_last": ([_fp, _fp, _fp, _fp, _i32, _i32, _fp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _fp],
                            C.c_int),
    "turtle_layernorm": ([_fp, _i32, _fp, _fp, _fp, _i32, _i64, _i32, _i32, _fp], C.c_int),
    "turtle_gemm": ([C.POINTER(GemmArgs), _fp], C.c_int),
    "turtle_dwconv3x3": ([_fp, _i32, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _fp],
                         C.c_int),
    "turtle_dwconv3x3_patch_rows": ([_fp, _i32, _fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_chan_gram": ([_fp, _i32, _i32, _fp, _i32, _i32, _i64, _i32, _i32, _i32, _fp, _fp, _fp, _i32, _fp],
                         C.c_int),
    "turtle_chan_softmax": ([_fp, _fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _fp, _fp, _fp], C.c_int),
    "turtle_chan_fold": ([_fp, _fp, _i32, _i32, _i32, _fp, _i32, _fp], C.c_int),
    "turtle_chan_gram_b": ([_fp, _i32, _i32, _i64, _fp, _i32, _i32, _i64, _i64, _i32, _i32, _i32, _fp, _fp, _fp, _i64, _i64,
                            _i32, _i32, _fp], C.c_int),
    "turtle_chan_softmax_b": ([_fp, _fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _fp, _fp, _i64, _i64, _i32, _fp], C.c_int),
    "turtle_chan_fold_b": ([_fp, _fp, _i32, _i32, _i32, _fp, _i32, _i32, _fp], C.c_int),
    "turtle_scale_cols": ([_fp, _i32, _i32, _fp, _fp, _i32, _i32, _i64, _i32, _i32, _fp], C.c_int),
    "turtle_sab_window_reduce": ([_fp, _i32, _fp, _fp, _fp, _i64, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_sab_window_reduce_h16": ([_fp, _i32, _fp, _fp, _fp, _i64, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_sab_patch_normalize": ([_fp, _i64, _i32, _fp], C.c_int),
    "turtle_sab_select": ([_fp, _fp, _i64, _i32, _i32, _i32, _i32, _fp, _i32, _fp, _fp, _i32, _fp], C.c_int),
    "turtle_sab_select_tc_workspace": ([_i32, _i32, _i32], C.c_longlong),
    "turtle_sab_select_tc": ([_fp, _fp, _i64, _i32, _i32, _i32, _i32, _fp, _i32, _fp, _fp, _fp, _fp], C.c_int),
    "turtle_sab_aggregate": ([_fp, _fp, _fp, _i64, _fp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_sab_aggregate_tc_workspace": ([_i32, _i32, _i32], C.c_longlong),
    "turtle_sab_aggregate_tc": ([_fp, _fp, _fp, _i32, _i64, _fp, _i32, _i32, _i32, _i32, _i32, _i32, _fp, _fp], C.c_int),
    "turtle_cast_f16": ([_fp, _fp, _i64, _fp], C.c_int),
    "turtle_add_posenc": ([_fp, _fp, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_ln2d_bwd_workspace": ([_i32, C.c_longlong], C.c_longlong),
    "turtle_ln2d_fwd": ([_fp, _i32, _fp, _fp, _fp, _fp, _fp, _i32, _i32, C.c_longlong, _fp], C.c_int),
    "turtle_ln2d_bwd": ([_fp, _fp, _i32, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i32, _i32, C.c_longlong, _fp], C.c_int),
    "turtle_ln2d_fwd_cast": ([_fp, _i32, _fp, _fp, _fp, _i32, _fp, _fp, _i32, _i32, C.c_longlong, _fp], C.c_int),
    "turtle_ln2d_bwd_cast": ([_fp, _i32, _fp, _i32, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i32, _i32, C.c_longlong, _fp],
                             C.c_int),
    "turtle_dwconv3x3_nchw": ([_fp, _i32, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_dwconv3x3_nchw_wgrad_workspace": ([_i32, _i32, _i32, _i32], C.c_longlong),
    "turtle_rownorm_fwd": ([_fp, _i32, _i64, _i32, _i32, _i32, _fp, _fp, _fp], C.c_int),
    "turtle_rownorm_bwd": ([_fp, _fp, _fp, _i32, _i64, _i32, _fp, _fp], C.c_int),
    "turtle_gelu_gate_nchw": ([_fp, _i32, _fp, _i32, _i32, _i64, _fp], C.c_int),
    "turtle_gelu_gate_nchw_bwd": ([_fp, _fp, _i32, _fp, _i32, _i32, _i64, _fp], C.c_int),
    "turtle_dwconv3x3_nchw_wgrad": ([_fp, _fp, _i32, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_gffw_fused": ([_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_gffw_tail": ([_fp, _fp, _fp, _fp, _fp, _fp, _fp, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_u8_to_frame": ([_fp, C.c_longlong, _fp, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_frame_to_u8": ([_fp, _fp, C.c_longlong, _i32, _i32, _i32, _i32, _i32, _fp], C.c_int),
    "turtle_frame_metrics_workspace": ([_i32, _i32], C.c_longlong),
    "turtle_frame_metrics": ([_fp, _fp, _i32, _i32, _i32, _i32, _fp, _fp, _fp], C.c_int),
    "turtle_tile_gather": ([_fp, _fp, _fp, _i32, _i32, _i32, _i32, C.POINTER(_i32), _i32, C.POINTER(_i32), _i32, _fp],
                           C.c_int),
    "turtle_tile_blend": ([_fp, _fp, _i32, _i32, _i32, _i32, C.POINTER(_i32), _i32, C.POINTER(_i32), _i32, _i32, _fp],
                          C.c_int),
    "turtle_grad_check_finite": ([_fp, _i64, _fp, _fp], C.c_int),
    "turtle_adamw_flat": ([_fp, _fp, _fp, _fp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _fp, _fp], C.c_int),
}

EXPORTS = tuple(_SIGS)
_lib = None
launch_count = 0          # number of kernel-launching C-ABI calls made by this process


class TurtleKernelError(RuntimeError):
    code = 0


def lib_path() -> str:
    """The in-tree library; TURTLE_LIB_PATH points at another build of the same ABI (A/B measurements)."""
    return os.environ.get("TURTLE_LIB_PATH") or _build.LIB


def load():
    """Load (building first if needed) libturtle_b200.so; raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        _build.build()
    lib = C.CDLL(path)
    for name, (argtypes, restype) in _SIGS.items():
        fn = getattr(lib, name)       # AttributeError if a declared symbol is not exported
        fn.argtypes, fn.restype = argtypes, restype
    if lib.turtle_sizeof_gemm_args() != C.sizeof(GemmArgs):
        raise TurtleKernelError("GemmArgs does not mirror TurtleGemmArgs of the loaded library (stale build?)")
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke a kernel-launching entry point; non-zero return codes raise."""
    global launch_count
    rc = getattr(load(), name)(*args)
    launch_count += 1
    if rc != 0:
        err = TurtleKernelError(f"{name} failed with code {rc}")
        err.code = rc
        raise err
