"""ctypes binding of libunimp_b200.so (the C ABI in include/unimp_b200.h).

There is no CPU path and no fallback: if the library is missing, importing the ops raises;
if a kernel returns non-zero, the call raises with unimp_last_error_string().
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunimp_b200.so")

F32, BF16 = 0, 1


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("batch_stride", C.c_int64), ("row_stride", C.c_int64)]


_i, _i64, _f, _p = C.c_int, C.c_int64, C.c_float, C.c_void_p

# name -> (restype, argtypes); mirrors include/unimp_b200.h one to one
SIGNATURES = {
    "unimp_version": (_i, []),
    "unimp_last_error_string": (C.c_char_p, []),
    "unimp_device_ok": (_i, []),
    "unimp_text_time": (_i, [_p, _i64, _i, _i, _i, _i, _p, _p]),
    "unimp_xattn_fwd": (_i, [View, View, View, _p, View, _p, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "unimp_attn_bwd_workspace": (_i64, [_i, _i, _i, _i, _i]),
    "unimp_xattn_bwd": (_i, [View, View, View, _p, View, View, _p, _p, View, View, View,
                             _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "unimp__attn_fwd_simt": (_i, [View, View, View, _p, View, _p, _i, _i, _i, _i, _i, _i, _i, _f,
                                  _i, _p]),
    "unimp__attn_bwd_simt": (_i, [View, View, View, _p, View, View, _p, _p, View, View, View,
                                  _i, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "unimp_attn_fwd": (_i, [View, View, View, View, _p, _i, _i, _i, _i, _i, _f, _i, _p]),
    "unimp_attn_bwd": (_i, [View, View, View, View, View, _p, _p, View, View, View,
                            _i, _i, _i, _i, _i, _f, _i, _p]),
    "unimp_xattn_block_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "unimp_xattn_block_fwd": (_i, [_p, _p, View, View, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i,
                                   _f, _i, _p]),
    "unimp_xattn_decode": (_i, [View, View, View, _p, View, _i, _i, _i, _i, _i, _f, _i, _p]),
    "unimp_lm_decode_attn": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _p]),
    "unimp_beam_topk_workspace": (_i64, [_i, _i]),
    "unimp_beam_topk": (_i, [_p, _i64, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "unimp_linear_small_m": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "unimp_gate_residual_ln_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _f, _i, _p]),
    "unimp_gate_residual_ln_bwd_workspace": (_i64, [_i64, _i]),
    "unimp_gate_residual_ln_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                        _i64, _i, _i, _i, _p]),
    "unimp_focal_ce_workspace": (_i64, [_i, _i, _i, _i]),
    "unimp_focal_ce_fwd": (_i, [_p, _i64, _p, _p, _f, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "unimp_focal_ce_bwd": (_i, [_p, _i64, _p, _p, _f, _i, _p, _p, _p, _p, _p, _i64, _i, _i, _i,
                                _i, _i, _p]),
    "unimp_focal_ce_rows_fwd": (_i, [_p, _i64, _p, _p, _p, _f, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "unimp_focal_ce_rows_bwd": (_i, [_p, _i64, _p, _p, _p, _f, _i, _p, _p, _p, _p, _p, _i64, _i, _i,
                                     _i, _i, _p]),
    "unimp_mask_labels": (_i, [_p, _i64, _i64, _i64, _i64, _p, _i, _i, _p]),
    "unimp_adamw_step": (_i, [_p, _p, _p, _p, _p, _i64, _p, _f, _f, _f, _f, _p, _f, _f, _i, _i, _p]),
    "unimp_sumsq": (_i, [_p, _i64, _p, _i, _p]),
    "unimp_rotary_qkv_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i, _p]),
    "unimp_rotary_qkv_bwd": (_i, [_p, _p, _p, C.POINTER(C.c_int64), _p, _p, _p, _i, _i, _i, _i, _i,
                                  _i64, _i, _p]),
    "unimp_lm_attn_supported": (_i, [_i, _i, _i, _i]),
    "unimp_key_bits": (_i, [_p, _i, _p, _i, _i, _p]),
    "unimp_lm_attn_fwd": (_i, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "unimp_lm_attn_bwd": (_i, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p,
                               _i, _i, _i, _i, _f, _i, _p]),
    "unimp_rotary_qkv_bwd_f32q": (_i, [_p, _p, _p, C.POINTER(C.c_int64), _p, _p, _p, _i, _i, _i, _i, _i,
                                       _i64, _i, _p]),
    "unimp_quick_gelu": (_i, [_p, _i64, _i, _p]),
    "unimp_gelu_fwd": (_i, [_p, _p, _i64, _i, _p]),
    "unimp_gelu_bwd": (_i, [_p, _p, _p, _i64, _i, _p]),
}

_lib = None


class UnimpError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UnimpError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or unimp_b200/csrc/build.sh). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().unimp_last_error_string().decode("utf-8", "replace")
        raise UnimpError(f"{what} failed (rc={rc}): {msg}")
