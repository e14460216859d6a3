"""ctypes binding of libs2f.so (include/s2f.h).  There is no fallback: a missing library is fatal."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S2F_LIB") or os.path.join(_HERE, "libs2f.so")      # S2F_LIB: experiment builds only

ABI_VERSION = 10


class ConvArgs(C.Structure):
    """s2f_conv_args (include/s2f.h)."""
    _fields_ = [
        ("a", C.c_void_p), ("a_is_spike", C.c_int), ("a_scale", C.c_float),
        ("w", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p),
        ("out_f32", C.c_void_p), ("out_spike", C.c_void_p), ("out_transposed", C.c_int),
        ("n", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
        ("KH", C.c_int), ("KW", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("a_img_stride", C.c_int64), ("a_stride_m", C.c_int64), ("a_stride_k", C.c_int64),
        ("w_img_stride", C.c_int64), ("d_max", C.c_float),
    ]


class GemmTcArgs(C.Structure):
    """s2f_gemm_tc_args (include/s2f.h)."""
    _fields_ = [
        ("a", C.c_void_p), ("w_packed", C.c_void_p),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p),
        ("out_f32", C.c_void_p), ("out_spike", C.c_void_p), ("out_transposed", C.c_int),
        ("n", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
        ("KH", C.c_int), ("KW", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("pieces", C.c_int),
        ("d_max", C.c_float), ("per_image_weights", C.c_int),
        ("up_prev", C.c_void_p), ("up_H", C.c_int), ("up_W", C.c_int),
        ("a_ld", C.c_int64),
    ]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes): every symbol include/s2f.h declares
SIGNATURES = {
    "s2f_last_error": (C.c_char_p, []),
    "s2f_abi_version": (_I, []),
    "s2f_launch_count": (C.c_uint64, []),
    "s2f_nilif_fwd": (_I, [_P, _P, _P, _P, _L, _P, _P, _P, _P, _I, _L, _I, _F, _F, _I, _I, _P, _P]),
    "s2f_nilif_pair": (_I, [_P, _P, _P, _P, _L, _P, _P, _L, _I, _F, _P]),
    "s2f_nilif_bwd": (_I, [_P, _P, _P, _P, _P, _P, _L, _I, _F, _F, _P]),
    "s2f_conv_simt": (_I, [C.POINTER(ConvArgs), _P]),
    "s2f_gemm_i8_tc": (_I, [C.POINTER(GemmTcArgs), _P]),
    "s2f_pack_weights_i8": (_L, [_P, _I, _I, _I, _I, _P, _P]),
    "s2f_pack_rows_i8_device": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _F, _P]),
    "s2f_dwconv": (_I, [_P, _I, _F, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "s2f_fpn_merge_f16": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P]),
    "s2f_sepconv_bpack_bytes": (_L, [_I, _I]),
    "s2f_sepconv_dwpw": (_I, [_P, _F, _P, _P, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "s2f_linear_attn_ws_bytes": (_L, [_I, _I, _I]),
    "s2f_linear_attn": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "s2f_affine_add_lif": (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _P]),
    "s2f_dcnv3_gather": (_I, [_P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "s2f_upsample_add_lif": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "s2f_sigmoid_lif": (_I, [_P, _P, _L, _F, _P]),
    "s2f_preprocess_u8": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _I, _F, _P]),
    "s2f_level_hist": (_I, [_P, _L, _P, _P]),
    "s2f_stem_u8": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "s2f_semantic_tail": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "s2f_semantic_tail_ws_bytes": (_L, [_I, _I]),
    "s2f_semantic_tail_tc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "s2f_peak_mma": (_L, [_I, _I, _P]),
    "s2f_nilif_train_fwd": (_I, [_P, _P, _P, _L, _F, _F, _P]),
    "s2f_nilif_train_bwd": (_I, [_P, _P, _P, _L, _F, _P]),
    "s2f_dec_attn": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
}

_lib = None


class S2FError(RuntimeError):
    pass


def lib():
    """Load libs2f.so once.  Raises (never falls back) when the library is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise S2FError(f"{LIB_PATH} is missing: build it with spike2former_b200/csrc/build.sh "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if handle.s2f_abi_version() != ABI_VERSION:
        raise S2FError(f"libs2f.so ABI {handle.s2f_abi_version()} != binding ABI {ABI_VERSION}: rebuild")
    _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().s2f_last_error().decode(errors="replace")
        raise S2FError(f"{what} failed (code {rc}): {msg}")
