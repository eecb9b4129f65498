"""ctypes binding of libfo1.so (the C ABI in include/fo1.h).  No CPU fallback: if the library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libfo1.so")

FO1_HFRE_MAX_LEVELS = 8
FO1_BF16, FO1_F32, FO1_I32, FO1_I64, FO1_U8, FO1_F16 = range(6)
EPI_NONE, EPI_GELU, EPI_SILU = range(3)


class Fo1Error(RuntimeError):
    pass


class HfreLevel(C.Structure):
    _fields_ = [("data", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("up_H", C.c_int32), ("up_W", C.c_int32), ("spatial_scale", C.c_float),
                ("box_set", C.c_int32), ("out_offset", C.c_int32)]


class HfreImage(C.Structure):
    _fields_ = [("levels", HfreLevel * FO1_HFRE_MAX_LEVELS), ("n_levels", C.c_int32), ("n_boxes", C.c_int32),
                ("boxes_aux", C.c_void_p), ("boxes_vt", C.c_void_p), ("out", C.c_void_p),
                ("out_bf16", C.c_void_p), ("pos_img_w", C.c_float), ("pos_img_h", C.c_float),
                ("pos_box_set", C.c_int32)]


class HfreParams(C.Structure):
    _fields_ = [("out_dim", C.c_int32), ("roi_size", C.c_int32), ("apply_pos_embed", C.c_int32),
                ("algo", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("A", C.c_void_p), ("lda", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64),
                ("D", C.c_void_p), ("ldd", C.c_int64), ("d_dtype", C.c_int32),
                ("bias", C.c_void_p), ("bias_dtype", C.c_int32), ("act", C.c_int32),
                ("residual", C.c_void_p), ("ldr", C.c_int64), ("gated", C.c_int32), ("tile_n", C.c_int32), ("split_k", C.c_int32)]


class AttnDesc(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
                ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64), ("ldo", C.c_int64),
                ("cu_seqlens", C.c_void_p), ("n_seqs", C.c_int32), ("max_seqlen", C.c_int32),
                ("q_heads", C.c_int32), ("kv_heads", C.c_int32), ("head_dim", C.c_int32),
                ("scale", C.c_float), ("causal", C.c_int32), ("total_rows", C.c_int32)]


class DecodeAttnDesc(C.Structure):
    _fields_ = [("q", C.c_void_p), ("ldq", C.c_int64), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p), ("cache_len", C.c_void_p),
                ("n_seqs", C.c_int32), ("cap", C.c_int32), ("q_heads", C.c_int32), ("kv_heads", C.c_int32), ("head_dim", C.c_int32),
                ("scale", C.c_float), ("out", C.c_void_p), ("ldo", C.c_int64), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


_lib = None


def lib() -> C.CDLL:
    """Load libfo1.so once.  Raises Fo1Error if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Fo1Error(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU or PyTorch fallback for the FO1 hot path)")
    L = C.CDLL(LIB_PATH)
    L.fo1_abi_version.restype = C.c_int
    L.fo1_last_error.restype = C.c_char_p
    L.fo1_launch_count.restype = C.c_uint64
    L.fo1_last_decode_path.restype = C.c_int
    L.fo1_last_decode_path.argtypes = [C.c_void_p]
    L.fo1_launch_count_reset.restype = None
    L.fo1_hfre_workspace_bytes.restype = C.c_size_t
    L.fo1_hfre_workspace_bytes.argtypes = [C.POINTER(HfreImage), C.c_int32, C.POINTER(HfreParams)]
    L.fo1_hfre_forward.restype = C.c_int
    L.fo1_hfre_forward.argtypes = [C.POINTER(HfreImage), C.c_int32, C.POINTER(HfreParams), C.c_void_p,
                                   C.c_size_t, C.c_void_p]
    L.fo1_decode_attention_workspace_bytes.restype = C.c_size_t
    L.fo1_decode_attention_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    L.fo1_decode_attention.restype = C.c_int
    L.fo1_decode_attention.argtypes = [C.POINTER(DecodeAttnDesc), C.c_void_p]
    L.fo1_gemm_bf16.restype = C.c_int
    L.fo1_gemm_bf16.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    L.fo1_attention_varlen.restype = C.c_int
    L.fo1_attention_varlen.argtypes = [C.POINTER(AttnDesc), C.c_void_p]
    _lib = L
    return L


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().fo1_last_error().decode("utf-8", "replace")
        raise Fo1Error(f"{what} failed with status {status}: {msg}")


def exported_symbols() -> list:
    """Every symbol include/fo1.h declares (used by the CPU test that the library loads)."""
    import re
    hdr = os.path.join(os.path.dirname(_PKG), "include", "fo1.h")
    text = open(hdr).read()
    return sorted(set(re.findall(r"\b(fo1_[a-z0-9_]+)\s*\(", text)))
