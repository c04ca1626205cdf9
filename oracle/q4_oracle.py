"""oracle/q4_oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy/ctypes front end of the CPU oracle (oracle/q4_oracle.c).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product (quantizations_b200/) never
does and has no CPU fallback.

Every function cites the reference (kkbwilldo/quantizations, /root/reference) file:line it restates.  Arrays are
numpy; 16-bit values travel as float32 holding the exactly-widened value (see q4_oracle.c header).

Parity pin: tests/golden/*.npz are outputs of the reference's own CUDA kernels run on a B200
(tests/golden/make_golden.py); tests/test_oracle_golden.py checks this module against them bit-for-bit.
NF4 quantize/dequantize: parity unpinned (no NF4 quantizer exists in the reference).
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libq4oracle.so")

F32, F16, BF16 = 0, 1, 2
FP4, NF4 = 1, 2
_QT = {"fp4": FP4, "nf4": NF4}
_DT = {"float32": F32, "float16": F16, "bfloat16": BF16, "fp32": F32, "fp16": F16, "bf16": BF16}

# sha256 of the little-endian float32 bytes of create_dynamic_map() as the reference builds it
DYNAMIC_MAP_SHA256 = "e732639a65f497b4ad684bb166a4467708255edd5207757de8b8f0c7e1fda89c"


def build(force: bool = False) -> str:
    """Compile oracle/q4_oracle.c with gcc (oracle/Makefile target `oracle`)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "q4_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i, l, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float
        L.q4o_quantize_fp4.restype = ctypes.c_ubyte
        L.q4o_quantize_fp4.argtypes = [f]
        L.q4o_quantize_nf4.restype = ctypes.c_ubyte
        L.q4o_quantize_nf4.argtypes = [f]
        L.q4o_dequantize_fp4.restype = f
        L.q4o_dequantize_fp4.argtypes = [ctypes.c_ubyte, f]
        L.q4o_dequantize_nf4.restype = f
        L.q4o_dequantize_nf4.argtypes = [ctypes.c_ubyte, f]
        L.q4o_quantize_8bit.restype = ctypes.c_ubyte
        L.q4o_quantize_8bit.argtypes = [vp, f]
        L.q4o_round_to.restype = f
        L.q4o_round_to.argtypes = [f, i]
        L.q4o_nf4_table.restype = ctypes.POINTER(ctypes.c_float * 16)
        L.q4o_num_threads.restype = i
        for name, args in {
            "q4o_quantize_blockwise_4bit": [vp, vp, vp, i, l, i],
            "q4o_quantize_blockwise_8bit": [vp, vp, vp, vp, i, l],
            "q4o_dequantize_blockwise_8bit": [vp, vp, vp, vp, i, l],
            "q4o_dequantize_absmax": [vp, vp, vp, f, vp, i, l],
            "q4o_dequantize_blockwise_4bit": [vp, vp, vp, i, l, i, i],
            "q4o_gemv_4bit": [vp, vp, vp, vp, vp, l, l, i, i, i],
            "q4o_gemv_4bit_f64": [vp, vp, vp, vp, vp, l, l, i],
            "q4o_linear_f64": [vp, vp, vp, l, l, l],
        }.items():
            fn = getattr(L, name)
            fn.restype = None
            fn.argtypes = args
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().q4o_num_threads())


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


# ----------------------------------------------------------------------------- code tables


def fp4_table() -> np.ndarray:
    """reference: core.py:193-229 (get_4bit_type("fp4")): [0,.0625,8,12,4,6,2,3,-0,...]/12 in float32.
    Index 8 is +0.0 here (Python's literal -0 is the integer 0) although the dequant tree yields -0.0."""
    data = np.array(
        [0, 0.0625, 8.0, 12.0, 4.0, 6.0, 2.0, 3.0, -0, -0.0625, -8.0, -12.0, -4.0, -6.0, -2.0, -3.0], dtype=np.float32
    )
    return (data / np.abs(data).max()).astype(np.float32)


def nf4_table() -> np.ndarray:
    """reference: csrc/kernels.cu:851 (q_data), the reference's only NF4 artefact."""
    return np.array(lib().q4o_nf4_table().contents, dtype=np.float32)


def code_table(quant_type: str) -> np.ndarray:
    return fp4_table() if quant_type == "fp4" else nf4_table()


_dynamic_map_cache: Optional[np.ndarray] = None


def dynamic_map() -> np.ndarray:
    """reference: core.py:251-314 (create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8)).

    The reference builds the table with torch.linspace on the CPU in float32; torch is third-party arithmetic on
    this path, so the restatement calls the same op (a numpy linspace rounds differently) and the result is pinned
    by DYNAMIC_MAP_SHA256.
    """
    global _dynamic_map_cache
    if _dynamic_map_cache is None:
        import torch

        exp_bits, nonsign = 7, 7
        vals = []
        for i in range(exp_bits):
            nfrac = int(2 ** (i + nonsign - exp_bits) + 1)
            edges = torch.linspace(0.1, 1, nfrac)
            centres = (edges[:-1] + edges[1:]) / 2.0
            scale = 10 ** (-(exp_bits - 1) + i)
            vals += (scale * centres).tolist()
            vals += (-scale * centres).tolist()
        # 2**(nonsign-exp_bits)-1 == 0 extra items for the default arguments (core.py:275,292)
        vals += [0, 1.0]
        assert len(vals) == 256
        vals.sort()
        _dynamic_map_cache = torch.Tensor(vals).numpy().astype(np.float32)
    return _dynamic_map_cache


def dynamic_map_sha256() -> str:
    return hashlib.sha256(dynamic_map().astype("<f4").tobytes()).hexdigest()


# ----------------------------------------------------------------------------- scalar codecs


def quantize_fp4_scalar(x: float) -> int:
    """reference: csrc/kernels.cu:113-163"""
    return int(lib().q4o_quantize_fp4(float(np.float32(x))))


def dequantize_fp4_scalar(v: int, absmax: float) -> np.float32:
    """reference: csrc/kernels.cu:70-111"""
    return np.float32(lib().q4o_dequantize_fp4(int(v), float(np.float32(absmax))))


def quantize_nf4_scalar(x: float) -> int:
    return int(lib().q4o_quantize_nf4(float(np.float32(x))))


def quantize_8bit_scalar(code: np.ndarray, x: float) -> int:
    """reference: csrc/kernels.cu:166-238"""
    code = _f32(code)
    return int(lib().q4o_quantize_8bit(_p(code), float(np.float32(x))))


def round_to(a: np.ndarray, dtype: str) -> np.ndarray:
    """float32 -> {fp16, bf16} round-to-nearest-even -> float32 (exact widening)."""
    dt = _DT[dtype]
    a = _f32(a)
    if dt == F32:
        return a.copy()
    if dt == F16:
        return a.astype(np.float16).astype(np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    r[nan] = 0x7FFF0000
    return r.view(np.float32)


# ----------------------------------------------------------------------------- blockwise ops


def quantize_blockwise_4bit(A: np.ndarray, blocksize: int = 64, quant_type: str = "fp4") -> Tuple[np.ndarray, np.ndarray]:
    """reference: csrc/kernels.cu:401-476 + core.py:536-559.  Returns (packed uint8[(n+1)//2], absmax float32[nb])."""
    A = _f32(A).reshape(-1)
    n = A.size
    nb = (n + blocksize - 1) // blocksize
    absmax = np.zeros(nb, dtype=np.float32)
    out = np.zeros((n + 1) // 2, dtype=np.uint8)
    lib().q4o_quantize_blockwise_4bit(_p(A), _p(absmax), _p(out), blocksize, n, _QT[quant_type])
    return out, absmax


def quantize_blockwise_8bit(A: np.ndarray, blocksize: int = 4096, code: Optional[np.ndarray] = None):
    """reference: core.py:317-366 + csrc/kernels.cu:396-461.  Returns (uint8[n], absmax float32[nb])."""
    code = dynamic_map() if code is None else _f32(code)
    A = _f32(A).reshape(-1)
    n = A.size
    nb = (n + blocksize - 1) // blocksize
    absmax = np.zeros(nb, dtype=np.float32)
    out = np.zeros(n, dtype=np.uint8)
    lib().q4o_quantize_blockwise_8bit(_p(code), _p(A), _p(absmax), _p(out), blocksize, n)
    return out, absmax


def dequantize_blockwise_8bit(q: np.ndarray, absmax: np.ndarray, blocksize: int = 4096, code: Optional[np.ndarray] = None):
    """reference: core.py:369-423 + csrc/kernels.cu:541,549-553."""
    code = dynamic_map() if code is None else _f32(code)
    q = _u8(q).reshape(-1)
    absmax = _f32(absmax)
    out = np.empty(q.size, dtype=np.float32)
    lib().q4o_dequantize_blockwise_8bit(_p(code), _p(q), _p(absmax), _p(out), blocksize, q.size)
    return out


def dequantize_absmax(qabsmax, absmax2, offset, blocksize2: int = 256, code2: Optional[np.ndarray] = None) -> np.ndarray:
    """reference: core.py:467-468 / :614-615 (dequantize_blockwise then `absmax += offset`)."""
    code2 = dynamic_map() if code2 is None else _f32(code2)
    q = _u8(qabsmax).reshape(-1)
    a2 = _f32(absmax2)
    out = np.empty(q.size, dtype=np.float32)
    lib().q4o_dequantize_absmax(_p(code2), _p(q), _p(a2), float(np.float32(offset)), _p(out), blocksize2, q.size)
    return out


def dequantize_blockwise_4bit(packed, absmax, n: int, blocksize: int = 64, quant_type: str = "fp4", out_dtype: str = "float16"):
    """reference: csrc/kernels.cu:528-567 + core.py:619-631.  Returns float32[n] holding values rounded to out_dtype."""
    packed = _u8(packed).reshape(-1)
    absmax = _f32(absmax)
    out = np.empty(n, dtype=np.float32)
    lib().q4o_dequantize_blockwise_4bit(_p(packed), _p(absmax), _p(out), blocksize, n, _QT[quant_type], _DT[out_dtype])
    return out


def quantize_4bit(A: np.ndarray, blocksize: int = 64, quant_type: str = "fp4", offset: Optional[float] = None,
                  compress_statistics: bool = True) -> dict:
    """reference: core.py:507-578 (quantize_4bit): 4-bit pack, offset = absmax.mean(), absmax -= offset, 8-bit
    blockwise quantize (blocksize 256) of the shifted absmax.

    `offset` is torch's CUDA reduction in the reference (core.py:563); its summation order is torch's, so parity
    tests pass the value the product computed with that same torch op.  With offset=None a float32 numpy mean is used
    (fine for self-consistent CPU use, not bit-comparable to a GPU mean).
    """
    packed, absmax = quantize_blockwise_4bit(A, blocksize, quant_type)
    st = {"packed": packed, "absmax_f32": absmax, "blocksize": blocksize, "quant_type": quant_type,
          "shape": tuple(np.shape(A)), "code": code_table(quant_type)}
    if compress_statistics:
        off = np.float32(absmax.mean(dtype=np.float32) if offset is None else offset)
        shifted = (absmax - off).astype(np.float32)
        qabsmax, absmax2 = quantize_blockwise_8bit(shifted, 256)
        st.update(offset=off, qabsmax=qabsmax, absmax2=absmax2, code2=dynamic_map())
    return st


def state_absmax(st: dict) -> np.ndarray:
    """fp32 per-block absmax the kernels consume (decoded from the double-quant state when present)."""
    if "qabsmax" in st:
        return dequantize_absmax(st["qabsmax"], st["absmax2"], st["offset"], 256, st["code2"])
    return st["absmax_f32"]


def dequantize_4bit(st: dict, out_dtype: str = "float16") -> np.ndarray:
    """reference: core.py:581-634 (without the trailing .t())."""
    n = int(np.prod(st["shape"]))
    return dequantize_blockwise_4bit(st["packed"], state_absmax(st), n, st["blocksize"], st["quant_type"], out_dtype).reshape(
        st["shape"]
    )


# ----------------------------------------------------------------------------- GEMV / linear


def gemv_4bit(x, packed, absmax, code, N: int, K: int, blocksize: int = 64, dtype: str = "float32", f32_fused: bool = True):
    """reference: csrc/kernels.cu:1126-1218 in the reference's own summation order (see q4_oracle.c)."""
    x = _f32(x).reshape(-1)
    packed = _u8(packed).reshape(-1)
    absmax = _f32(absmax)
    code = _f32(code)
    out = np.empty(N, dtype=np.float32)
    lib().q4o_gemv_4bit(_p(x), _p(packed), _p(absmax), _p(code), _p(out), N, K, blocksize, _DT[dtype], int(f32_fused))
    return out


def gemv_4bit_f64(x, packed, absmax, code, N: int, K: int, blocksize: int = 64) -> np.ndarray:
    x = _f32(x).reshape(-1)
    packed = _u8(packed).reshape(-1)
    absmax = _f32(absmax)
    code = _f32(code)
    out = np.empty(N, dtype=np.float64)
    lib().q4o_gemv_4bit_f64(_p(x), _p(packed), _p(absmax), _p(code), _p(out), N, K, blocksize)
    return out


def linear_f64(X, Wdeq) -> np.ndarray:
    """reference: modules.py:63-64 (F.linear(A, dequantize_4bit(B).t())) in fp64."""
    X = _f32(X)
    W = _f32(Wdeq)
    M, K = X.reshape(-1, X.shape[-1]).shape
    N = W.shape[0]
    out = np.empty((M, N), dtype=np.float64)
    lib().q4o_linear_f64(_p(X), _p(W), _p(out), M, N, K)
    return out
