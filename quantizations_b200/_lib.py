"""ctypes binding of libquantizations_b200.so (the C ABI in include/quantizations_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this module raises.  The library is built
in-tree by `python -m quantizations_b200.build` (also run by __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("Q4_LIB_PATH") or os.path.join(_HERE, "libquantizations_b200.so")  # override: developer A/B builds

Q4_F32, Q4_F16, Q4_BF16 = 0, 1, 2
Q4_GENERAL8BIT, Q4_FP4, Q4_NF4 = 0, 1, 2
Q4_GEMV_DEFAULT, Q4_GEMV_EXACT_F32, Q4_GEMV_PDL, Q4_GEMV_SHARE_SM, Q4_ATTN_EARLY_CACHE, Q4_GEMV_SWIGLU = 0, 1, 2, 4, 8, 16
Q4_GEMV_BATCH_TC5 = 32
Q4_ERR_SHAPE, Q4_ERR_ALIGN = -4, -6  # include/quantizations_b200.h: q4_status
Q4_GEMV_RING_MAX_STAGES, Q4_GEMV_RING_WS_BYTES = 32, 73728 + 8 * 131072
Q4_GEMV_LUT_BYTES = 65536
Q4_GEMV_WORKSPACE_BYTES = 8 << 20


class Q4Error(RuntimeError):
    pass


class AbsmaxStats(ctypes.Structure):
    """q4_absmax_t"""

    _fields_ = [
        ("absmax", ctypes.c_void_p),
        ("qabsmax", ctypes.c_void_p),
        ("code2", ctypes.c_void_p),
        ("absmax2", ctypes.c_void_p),
        ("offset", ctypes.c_void_p),
        ("blocksize2", ctypes.c_int),
    ]


class AllReduce(ctypes.Structure):
    """q4_allreduce_t"""

    _fields_ = [("peer_bases", ctypes.c_void_p), ("world", ctypes.c_int), ("rank", ctypes.c_int), ("max_rows", ctypes.c_int)]


Q4_AR_MAX_CTAS = 1024
Q4_AR_DATA_OFFSET = 65536


def ar_bytes(max_rows: int) -> int:
    return Q4_AR_DATA_OFFSET + 2 * 8 * max_rows * 8


class GemvFused(ctypes.Structure):
    """q4_gemv_fused_t"""

    _fields_ = [
        ("x", ctypes.c_void_p), ("x_gate", ctypes.c_void_p), ("rms_weight", ctypes.c_void_p), ("rms_eps", ctypes.c_float),
        ("B", ctypes.c_void_p), ("stats", ctypes.POINTER(AbsmaxStats)), ("offsets", ctypes.POINTER(ctypes.c_void_p)),
        ("row_end", ctypes.POINTER(ctypes.c_int)), ("nmat", ctypes.c_int), ("code", ctypes.c_void_p), ("bias", ctypes.c_void_p),
        ("out", ctypes.c_void_p), ("rows", ctypes.c_int64), ("K", ctypes.c_int64), ("blocksize", ctypes.c_int),
        ("dtype", ctypes.c_int), ("flags", ctypes.c_int), ("prefetch", ctypes.c_void_p), ("prefetch_bytes", ctypes.c_int64),
        ("lut", ctypes.c_void_p), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64),
        ("allreduce", ctypes.POINTER(AllReduce)), ("prefetch_K", ctypes.c_int64),
    ]  # fmt: skip


_vp, _i, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
_SIGNATURES = {
    # reference names (pythonInterface.cpp:154-161)
    "cgemm_4bit_inference_naive_fp32": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
    "cquantize_blockwise_fp16_fp4": [_vp, _vp, _vp, _vp, _i, _i],
    "cdequantize_blockwise_fp16_fp4": [_vp, _vp, _vp, _vp, _i, _i],
    "cquantize_blockwise_fp32": [_vp, _vp, _vp, _vp, _i, _i],
    "cdequantize_blockwise_fp32": [_vp, _vp, _vp, _vp, _i, _i],
    # generalised API
    "q4_quantize_blockwise_4bit": [_vp, _vp, _vp, _i, _i64, _i, _i, _vp],
    "q4_quantize_blockwise_8bit": [_vp, _vp, _vp, _vp, _i, _i64, _vp],
    "q4_dequantize_blockwise_8bit": [_vp, _vp, _vp, _vp, _i, _i64, _vp],
    "q4_dequantize_blockwise_4bit": [_vp, ctypes.POINTER(AbsmaxStats), _vp, _i, _i64, _i, _i, _vp],
    "q4_gemv_4bit_grouped": [_vp, _vp, ctypes.POINTER(AbsmaxStats), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int), _i,
                             _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _vp, _i64, _vp],
    "q4_gemv_4bit_fused": [ctypes.POINTER(GemvFused), _vp],
    "q4_gemv_lut_build": [_vp, _vp, _i, _vp, _vp],
    "q4_gemv_4bit_chain": [ctypes.POINTER(GemvFused), _i, _vp, _vp],
    "q4_gemv_4bit_ring": [ctypes.POINTER(GemvFused), _i, _vp, _i64, _vp],
    "q4_gemv_4bit_batch": [_vp, _vp, ctypes.POINTER(AbsmaxStats), _vp, _vp, _vp, _i, _i64, _i64, _i, _i, _i, _vp, _vp, _i64, _vp],
    "q4_decode_attention": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "q4_argmax": [_vp, _i64, _i, _vp, _vp, _vp],
    "q4_gemm_4bit": [_vp, _vp, ctypes.POINTER(AbsmaxStats), _vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _vp],
    "q4_gemv_4bit": [_vp, _vp, ctypes.POINTER(AbsmaxStats), _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _vp, _i64, _vp],
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly if it is not built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Q4Error(
                f"{LIB_PATH} not found: build it with `python -m quantizations_b200.build` "
                "(quantizations_b200 has no CPU or PyTorch fallback)"
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library is stale: also loud
            fn.argtypes = args
            fn.restype = ctypes.c_int
        L.q4_abi_version.restype = ctypes.c_int
        L.q4_error_string.restype = ctypes.c_char_p
        L.q4_error_string.argtypes = [ctypes.c_int]
        L.q4_launch_count.restype = ctypes.c_int64
        L.q4_device_info.argtypes = [ctypes.POINTER(ctypes.c_int)] * 3
        L.q4_device_info.restype = ctypes.c_int
        _lib = L
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        raise Q4Error(f"{what} failed: {lib().q4_error_string(code).decode()} (code {code})")


def launch_count() -> int:
    return int(lib().q4_launch_count())


def exported_symbols():
    return list(_SIGNATURES) + ["q4_abi_version", "q4_error_string", "q4_launch_count", "q4_device_info"]
