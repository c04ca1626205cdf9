"""Generate tests/golden/*.npz from the REFERENCE's own CUDA kernels.  TEST INFRASTRUCTURE ONLY.

Run on a B200 (no GPU in the build container):

    make -C oracle ref                      # here: builds oracle/_ref/{kbkim_lib,ref_shim}.so from /root/reference
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
    cp gpurun_out/golden/*.npz tests/golden/

What is executed is the reference's unmodified object code (csrc/kernels.cu + csrc/ops.cu compiled for sm_100a):
  * the five functions the reference exports, through its own CPython module `kbkim_lib` (pythonInterface.cpp);
  * template instances it instantiates but does not export (bf16/fp32 quantize+dequantize, fp16/bf16 GEMV), through
    oracle/ref_shim.cu, which only forwards to the reference's launchers.
Inputs come from numpy's PCG64 with fixed seeds, so the fixtures are reproducible; every .npz holds the inputs AND the
reference's outputs, so the CPU oracle (oracle/q4_oracle.py) can be pinned against them without a GPU, and the product
kernels can be compared with them on a GPU without the reference being present.
"""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import q4_oracle as orc  # noqa: E402  (tables only: dynamic map, fp4/nf4 code)

import kbkim_lib  # noqa: E402  the reference's CPython module

shim = ctypes.CDLL(os.path.join(REF, "ref_shim.so"))
for _n in dir(shim):
    pass
_vp, _i = ctypes.c_void_p, ctypes.c_int
for name, args in {
    "ref_gemv_fp32": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
    "ref_gemv_fp16": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
    "ref_gemv_bf16": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
    "ref_quant_fp4_fp16": [_vp, _vp, _vp, _i, _i],
    "ref_quant_fp4_bf16": [_vp, _vp, _vp, _i, _i],
    "ref_quant_fp4_fp32": [_vp, _vp, _vp, _i, _i],
    "ref_quant_8bit_fp32": [_vp, _vp, _vp, _vp, _i, _i],
    "ref_dequant_fp4_fp16": [_vp, _vp, _vp, _i, _i],
    "ref_dequant_fp4_bf16": [_vp, _vp, _vp, _i, _i],
    "ref_dequant_fp4_fp32": [_vp, _vp, _vp, _i, _i],
    "ref_dequant_8bit_fp32": [_vp, _vp, _vp, _vp, _i, _i],
}.items():
    getattr(shim, name).argtypes = args
    getattr(shim, name).restype = _i

DEV = torch.device("cuda:0")
TDT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


def dev(a: np.ndarray, dtype=None) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def host_f32(t: torch.Tensor) -> np.ndarray:
    return t.float().cpu().numpy()


def ok(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: cuda error {rc}")


def sync():
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ input generators


def weights(rng, n, kind):
    if kind == "normal":
        return (rng.standard_normal(n) * 0.02).astype(np.float32)
    if kind == "uniform":
        return rng.uniform(-1, 1, n).astype(np.float32)
    if kind == "adversarial":
        a = (rng.standard_normal(n) * 0.02).astype(np.float32)
        if n >= 64 * 6:
            a[0:64] = 0.0  # all-zero block -> absmax 0 -> inv = inf -> NaN -> code 0
            a[64:128] = rng.choice([0.0, -0.0], 64).astype(np.float32)
            a[130] = 1e4  # single outlier
            a[192:256] = rng.uniform(-1, 1, 64).astype(np.float32) * 1e-7  # tiny values (fp16 subnormal / flush range)
            # a block whose normalised values sit exactly on / next to the FP4 thresholds
            thr = np.array([0.00260417, 0.0859375, 0.20833333, 0.29166667, 0.4166667, 0.583333, 0.8333333], dtype=np.float32)
            blk = np.concatenate([[1.0], thr, np.nextafter(thr, np.float32(2)), np.nextafter(thr, np.float32(0)), -thr])
            a[256 : 256 + blk.size] = blk.astype(np.float32)
            a[320:384] = -np.abs(a[320:384])  # all-negative block
        return a
    raise ValueError(kind)


def as_dtype_exact(a: np.ndarray, dtype: str) -> np.ndarray:
    """round to dtype and widen back to float32 (what the device tensor will hold)"""
    return host_f32(torch.from_numpy(a).to(TDT[dtype]))


# ------------------------------------------------------------------------------------------------ cases


def run_quant4(a_f32: np.ndarray, dtype: str, blocksize: int):
    n = a_f32.size
    A = dev(a_f32, TDT[dtype])
    nb = -(n // -blocksize)
    absmax = torch.zeros(nb, device=DEV, dtype=torch.float32)
    out = torch.zeros((n + 1) // 2, device=DEV, dtype=torch.uint8)
    if dtype == "float16":
        kbkim_lib.cquantize_blockwise_fp16_fp4(0, A.data_ptr(), absmax.data_ptr(), out.data_ptr(), blocksize, n)
        sync()
    else:
        fn = shim.ref_quant_fp4_bf16 if dtype == "bfloat16" else shim.ref_quant_fp4_fp32
        ok(fn(A.data_ptr(), absmax.data_ptr(), out.data_ptr(), blocksize, n), "quant")
    return out.cpu().numpy(), absmax.cpu().numpy()


def run_quant8(code: np.ndarray, a: np.ndarray, blocksize: int):
    n = a.size
    A, C = dev(a), dev(code)
    nb = -(n // -blocksize)
    absmax = torch.zeros(nb, device=DEV, dtype=torch.float32)
    out = torch.zeros(n, device=DEV, dtype=torch.uint8)
    kbkim_lib.cquantize_blockwise_fp32(C.data_ptr(), A.data_ptr(), absmax.data_ptr(), out.data_ptr(), blocksize, n)
    sync()
    return out.cpu().numpy(), absmax.cpu().numpy()


def run_dequant4(packed: np.ndarray, absmax: np.ndarray, n: int, dtype: str, blocksize: int):
    P, AM = dev(packed), dev(absmax)
    out = torch.zeros(n, device=DEV, dtype=TDT[dtype])
    if dtype == "float16":
        kbkim_lib.cdequantize_blockwise_fp16_fp4(0, P.data_ptr(), AM.data_ptr(), out.data_ptr(), blocksize, n)
        sync()
    else:
        fn = shim.ref_dequant_fp4_bf16 if dtype == "bfloat16" else shim.ref_dequant_fp4_fp32
        ok(fn(P.data_ptr(), AM.data_ptr(), out.data_ptr(), blocksize, n), "dequant")
    return host_f32(out)


def run_dequant8(code: np.ndarray, q: np.ndarray, absmax: np.ndarray, blocksize: int):
    C, Q, AM = dev(code), dev(q), dev(absmax)
    out = torch.zeros(q.size, device=DEV, dtype=torch.float32)
    kbkim_lib.cdequantize_blockwise_fp32(C.data_ptr(), Q.data_ptr(), AM.data_ptr(), out.data_ptr(), blocksize, q.size)
    sync()
    return out.cpu().numpy()


def run_gemv(x_f32, packed, absmax, code, N, K, dtype, blocksize):
    X = dev(x_f32, TDT[dtype])
    P, AM, C = dev(packed), dev(absmax), dev(code)
    out = torch.zeros(N, device=DEV, dtype=TDT[dtype])
    args = (N, 1, K, X.data_ptr(), P.data_ptr(), AM.data_ptr(), C.data_ptr(), out.data_ptr(), N, (K + 1) // 2, N, blocksize)
    if dtype == "float32":
        kbkim_lib.cgemm_4bit_inference_naive_fp32(*args)
        sync()
    else:
        ok((shim.ref_gemv_fp16 if dtype == "float16" else shim.ref_gemv_bf16)(*args), "gemv")
    return host_f32(out)


def main(outdir: str):
    os.makedirs(outdir, exist_ok=True)
    print("device:", torch.cuda.get_device_name(0))
    dyn = orc.dynamic_map()
    fp4, nf4 = orc.fp4_table(), orc.nf4_table()

    # ---- 1. 4-bit quantize (FP4): dtype x blocksize x size x distribution
    cases = {}
    rng = np.random.default_rng(20240501)
    idx = 0
    for dtype in ("float16", "bfloat16", "float32"):
        for blocksize, sizes in ((64, (64 * 64, 1000, 1001, 63, 1)), (128, (2048 + 37,)), (256, (4096,)), (512, (4096 + 5,)),
                                 (1024, (8192,)), (2048, (8192 + 100,)), (4096, (16384,))):
            for n in sizes:
                for kind in ("normal", "uniform", "adversarial"):
                    if kind != "normal" and (blocksize != 64 or n != 64 * 64):
                        continue
                    a = as_dtype_exact(weights(rng, n, kind), dtype)
                    packed, absmax = run_quant4(a, dtype, blocksize)
                    k = f"c{idx:03d}"
                    cases[k + "_meta"] = np.array([dtype, str(blocksize), str(n), kind])
                    cases[k + "_in"] = a
                    cases[k + "_packed"] = packed
                    cases[k + "_absmax"] = absmax
                    idx += 1
    np.savez_compressed(os.path.join(outdir, "quantize_fp4.npz"), **cases)
    print("quantize_fp4:", idx, "cases")

    # ---- 2. 8-bit codebook quantize of fp32 (double-quant of absmax): shifted absmax-like data, uniform, exact ties
    cases = {"code": dyn}
    rng = np.random.default_rng(20240502)
    mids = ((dyn[1:] + dyn[:-1]) * np.float32(0.5)).astype(np.float32)
    tie_block = np.concatenate([[1.0], dyn, mids, np.nextafter(mids, np.float32(2)), np.nextafter(mids, np.float32(-2))]).astype(np.float32)
    inputs = {
        "absmax_like": (np.abs(rng.standard_normal(256 * 40 + 17)) * 0.01 + 0.06).astype(np.float32),
        "uniform": rng.uniform(-1, 1, 4096 * 3 + 123).astype(np.float32),
        "ties": np.resize(tie_block, 4096).astype(np.float32),
        "zeros": np.zeros(300, dtype=np.float32),
    }
    inputs["absmax_like"] = (inputs["absmax_like"] - inputs["absmax_like"].mean(dtype=np.float32)).astype(np.float32)
    idx = 0
    for name, a in inputs.items():
        for blocksize in (256, 4096, 64):
            q, am = run_quant8(dyn, a, blocksize)
            k = f"c{idx:03d}"
            cases[k + "_meta"] = np.array([name, str(blocksize)])
            cases[k + "_in"] = a
            cases[k + "_q"] = q
            cases[k + "_absmax"] = am
            cases[k + "_deq"] = run_dequant8(dyn, q, am, blocksize)
            idx += 1
    np.savez_compressed(os.path.join(outdir, "blockwise_8bit.npz"), **cases)
    print("blockwise_8bit:", idx, "cases")

    # ---- 3. 4-bit dequantize (FP4): every byte value x assorted absmax, per dtype and blocksize; ragged / odd n
    cases = {}
    rng = np.random.default_rng(20240503)
    idx = 0
    for dtype in ("float16", "bfloat16", "float32"):
        for blocksize, n in ((64, 64 * 256), (64, 1001), (64, 7), (128, 4096 + 3), (4096, 8192 + 64)):
            nbytes = (n + 1) // 2
            packed = rng.integers(0, 256, nbytes, dtype=np.uint8)
            packed[: min(256, nbytes)] = np.arange(min(256, nbytes), dtype=np.uint8)
            nb = -(n // -blocksize)
            absmax = np.abs(rng.standard_normal(nb)).astype(np.float32) * 0.05 + 1e-3
            absmax[0] = 1.0
            if nb > 4:
                absmax[1], absmax[2], absmax[3] = 0.0, 3.0e-6, 70000.0  # zero, tiny (fp16 subnormal outputs), overflow in fp16
            out = run_dequant4(packed, absmax, n, dtype, blocksize)
            k = f"c{idx:03d}"
            cases[k + "_meta"] = np.array([dtype, str(blocksize), str(n)])
            cases[k + "_packed"] = packed
            cases[k + "_absmax"] = absmax
            cases[k + "_out"] = out
            idx += 1
    np.savez_compressed(os.path.join(outdir, "dequantize_fp4.npz"), **cases)
    print("dequantize_fp4:", idx, "cases")

    # ---- 4. GEMV: fp32 (exported), fp16/bf16 (shim); FP4 and NF4 code tables; K not a multiple of 1024; tall/skinny
    cases = {}
    rng = np.random.default_rng(20240504)
    idx = 0
    for dtype in ("float32", "float16", "bfloat16"):
        for (N, K) in ((64, 256), (33, 1088), (16, 4096), (128, 64), (8, 14336)):
            for code_name, code in (("fp4", fp4), ("nf4", nf4)):
                n = N * K
                packed = rng.integers(0, 256, n // 2, dtype=np.uint8)
                absmax = (np.abs(rng.standard_normal(n // 64)) * 0.01 + 0.05).astype(np.float32)
                x = as_dtype_exact(rng.standard_normal(K).astype(np.float32), dtype)
                out = run_gemv(x, packed, absmax, code, N, K, dtype, 64)
                k = f"c{idx:03d}"
                cases[k + "_meta"] = np.array([dtype, str(N), str(K), code_name])
                cases[k + "_x"] = x
                cases[k + "_packed"] = packed
                cases[k + "_absmax"] = absmax
                cases[k + "_code"] = code
                cases[k + "_out"] = out
                idx += 1
    np.savez_compressed(os.path.join(outdir, "gemv.npz"), **cases)
    print("gemv:", idx, "cases")

    # ---- 5. the reference's whole quantize_4bit recipe (core.py:536-576) on one Linear-shaped weight, step by step,
    #         including torch's CUDA mean for the offset, then its gemv_4bit recipe (core.py:467-499)
    rng = np.random.default_rng(20240505)
    N, K = 128, 512
    w = as_dtype_exact((rng.standard_normal(N * K) * 0.02).astype(np.float32), "float16")
    packed, absmax = run_quant4(w, "float16", 64)
    am_t = dev(absmax)
    offset_t = am_t.mean()
    shifted_t = am_t - offset_t
    shifted = shifted_t.cpu().numpy()
    qabs, absmax2 = run_quant8(dyn, shifted, 256)
    deq_abs = run_dequant8(dyn, qabs, absmax2, 256)
    deq_abs_t = dev(deq_abs)
    deq_abs_t += offset_t
    absmax_rt = deq_abs_t.cpu().numpy()
    wdeq = run_dequant4(packed, absmax_rt, N * K, "float16", 64)
    x = rng.standard_normal(K).astype(np.float32)
    y = run_gemv(x, packed, absmax_rt, fp4, N, K, "float32", 64)
    np.savez_compressed(
        os.path.join(outdir, "linear_fp4_recipe.npz"),
        w=w, packed=packed, absmax=absmax, offset=np.float32(offset_t.item()), shifted=shifted, qabsmax=qabs,
        absmax2=absmax2, absmax_roundtrip=absmax_rt, wdeq=wdeq, x=x, y=y, shape=np.array([N, K]),
    )
    print("linear_fp4_recipe: done")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
