"""Live-reference parity at BASELINE.json's full sizes (the four Llama-3-8B Linear shapes): the product's CUDA path against the
REFERENCE's own object code -- oracle/_ref/kbkim_lib.so (the unmodified CPython module, pythonInterface.cpp:154-178) and
oracle/_ref/ref_shim.so (forwarders to template instances the reference instantiates but does not export, ops.cu:175-197) --
executed on the same device, on identical seeded random-init weights.

  * FP4 quantize (fp16 through the exported entry point, bf16 through the shim): packed bytes and absmax BIT-EXACT
  * FP4 dequantize (fp16 / bf16): BIT-EXACT
  * GEMV fp32 (exported) and bf16 (shim), FP4 *and* NF4 code tables: max|y - y_ref| <= 1e-2 * max|y_ref| (north_star tolerance),
    and both within 1e-2 of an fp64 product of the dequantised weight
  * NF4 dequantize pinned to the reference: the reference GEMV is codebook-agnostic (kernels.cu:1119-1120; table at :851), so its
    exported fp32 instance with a one-hot activation returns code[nib] * absmax exactly -- compared BIT-FOR-BIT with
    dequantize_4bit(..., nf4) in fp32
  * NF4 quantize (no reference quantiser exists, ops.cuh:6-10): the property that pins it to the pinned table -- every stored
    nibble is a nearest entry of the table to the normalised value

Skipped only when oracle/_ref was not built (no /root/reference at build time).
"""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
pytestmark = pytest.mark.gpu

DEV = "cuda:0"
SHAPES = [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336)]  # q/o, k/v, gate/up, down (BASELINE configs[1])
TDT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


@pytest.fixture(scope="module")
def ref():
    """(kbkim_lib, shim): the reference's compiled code, loaded into this process"""
    if not (os.path.exists(os.path.join(REF, "kbkim_lib.so")) and os.path.exists(os.path.join(REF, "ref_shim.so"))):
        pytest.skip("oracle/_ref not built (make -C oracle ref needs /root/reference)")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kbkim_lib

    shim = ctypes.CDLL(os.path.join(REF, "ref_shim.so"))
    vp, i = ctypes.c_void_p, ctypes.c_int
    for name, args in {
        "ref_gemv_fp32": [i, i, i, vp, vp, vp, vp, vp, i, i, i, i], "ref_gemv_fp16": [i, i, i, vp, vp, vp, vp, vp, i, i, i, i],
        "ref_gemv_bf16": [i, i, i, vp, vp, vp, vp, vp, i, i, i, i], "ref_quant_fp4_bf16": [vp, vp, vp, i, i],
        "ref_dequant_fp4_bf16": [vp, vp, vp, i, i], "ref_dequant_fp4_fp32": [vp, vp, vp, i, i],
    }.items():  # fmt: skip
        getattr(shim, name).argtypes = args
        getattr(shim, name).restype = i
    return kbkim_lib, shim


@pytest.fixture(scope="module")
def q():
    import quantizations_b200 as q

    q._lib.lib()
    return q


def weight(N, K, dtype, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(N, K, device=DEV, dtype=torch.float32, generator=g) * 0.02).to(dtype)  # HF initializer_range


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_fp4_quantize_and_dequantize_bit_exact_vs_live_reference(q, ref, N, K, dtype):
    kbkim_lib, shim = ref
    n = N * K
    W = weight(N, K, TDT[dtype])
    absmax_r = torch.zeros(n // 64, device=DEV, dtype=torch.float32)
    packed_r = torch.zeros(n // 2, device=DEV, dtype=torch.uint8)
    if dtype == "float16":  # the reference's exported entry point (core.py:552)
        kbkim_lib.cquantize_blockwise_fp16_fp4(0, W.data_ptr(), absmax_r.data_ptr(), packed_r.data_ptr(), 64, n)
        torch.cuda.synchronize()
    else:
        assert shim.ref_quant_fp4_bf16(W.data_ptr(), absmax_r.data_ptr(), packed_r.data_ptr(), 64, n) == 0
    packed, state = q.quantize_4bit(W, quant_type="fp4", compress_statistics=False)
    assert torch.equal(packed.view(-1), packed_r), "packed bytes differ from the reference kernel's"
    assert torch.equal(state.absmax.view(torch.int32), absmax_r.view(torch.int32)), "absmax differs from the reference kernel's"
    # dequantize: the reference's kernel on the reference's own quantisation
    out_r = torch.zeros(n, device=DEV, dtype=TDT[dtype])
    if dtype == "float16":
        kbkim_lib.cdequantize_blockwise_fp16_fp4(0, packed_r.data_ptr(), absmax_r.data_ptr(), out_r.data_ptr(), 64, n)
        torch.cuda.synchronize()
    else:
        assert shim.ref_dequant_fp4_bf16(packed_r.data_ptr(), absmax_r.data_ptr(), out_r.data_ptr(), 64, n) == 0
    out = q.dequantize_4bit(packed, state).t().contiguous().view(-1)
    assert out.dtype == TDT[dtype]
    assert torch.equal(out.view(torch.int16), out_r.view(torch.int16)), "dequantised weight differs from the reference kernel's"


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("quant_type", ["fp4", "nf4"])
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_gemv_within_tolerance_of_live_reference(q, ref, N, K, quant_type, dtype):
    """The reference's gemv_4bit call sequence (core.py:467-499) with OUR decoded absmax (bit-exact with the reference's, see the
    golden tests) fed to ITS kernel; fp32 = the only instance it exports, bf16 = the instance it instantiates (ops.cu:177)."""
    kbkim_lib, shim = ref
    W = weight(N, K, torch.bfloat16, seed=1)
    packed, state = q.quantize_4bit(W, quant_type=quant_type)  # nested statistics, as every BASELINE config has them
    absmax = q.dequantize_blockwise(state.absmax, state.state2) + state.offset  # core.py:467-468
    x = torch.randn(1, 1, K, device=DEV, dtype=TDT[dtype], generator=torch.Generator(device=DEV).manual_seed(2))
    y_r = torch.zeros(N, device=DEV, dtype=TDT[dtype])
    args = (N, 1, K, x.data_ptr(), packed.data_ptr(), absmax.data_ptr(), state.code.data_ptr(), y_r.data_ptr(), N, (K + 1) // 2, N, 64)
    if dtype == "float32":
        kbkim_lib.cgemm_4bit_inference_naive_fp32(*args)
        torch.cuda.synchronize()
    else:
        assert shim.ref_gemv_bf16(*args) == 0
    y = q.gemv_4bit(x, packed, state=state).view(-1)
    # truth: fp64 product of the weight dequantised in fp32 (code * absmax, NOT rounded to the storage dtype of the state)
    st32 = q.QuantState(absmax=absmax, shape=state.shape, code=state.code, blocksize=64, quant_type=quant_type, dtype=torch.float32)
    truth = q.dequantize_4bit(packed, st32).t().double() @ x.view(-1).double()
    scale = truth.abs().max().item()
    err_vs_ref = (y.double() - y_r.double()).abs().max().item() / y_r.double().abs().max().item()
    err = (y.double() - truth).abs().max().item() / scale
    err_ref = (y_r.double() - truth).abs().max().item() / scale
    tol = 1e-2  # BASELINE north_star: "max rel err <= 1e-2 with fp32 accumulation"
    assert err_vs_ref <= tol, f"ours vs reference kernel: {err_vs_ref}"
    assert err <= (1e-5 if dtype == "float32" else 8e-3), f"ours vs truth: {err} (reference kernel: {err_ref})"
    assert err_ref <= tol


@pytest.mark.parametrize("N,K", [(4096, 4096), (1024, 4096), (512, 14336)])
def test_nf4_dequantize_pinned_by_the_reference_gemv(q, ref, N, K):
    """One-hot activations through the reference's exported fp32 GEMV with the NF4 table (kernels.cu:851 values) return
    code[nib] * absmax (fp32 multiply, kernels.cu:1166-1170) plus exact zeros: the reference's own arithmetic for an NF4
    dequantise.  dequantize_4bit(nf4) in fp32 must reproduce it bit-for-bit; the 16-bit outputs are its correctly rounded
    values."""
    kbkim_lib, _ = ref
    W = weight(N, K, torch.float32, seed=3)
    W[5, :64] = 0.0  # an all-zero block: absmax 0
    packed, state = q.quantize_4bit(W, quant_type="nf4", compress_statistics=False)
    assert set(np.unique(packed.cpu().numpy() >> 4)) | set(np.unique(packed.cpu().numpy() & 15)) == set(range(16))
    ours = q.dequantize_4bit(packed, state).t().contiguous()  # fp32 [N, K]
    assert ours.dtype == torch.float32
    cols = sorted({0, 1, 63, 64, 65, K // 2 - 1, K // 2, K - 2, K - 1} | set(np.random.default_rng(0).integers(0, K, 40).tolist()))
    y = torch.zeros(N, device=DEV, dtype=torch.float32)
    for k in cols:
        x = torch.zeros(K, device=DEV, dtype=torch.float32)
        x[k] = 1.0
        kbkim_lib.cgemm_4bit_inference_naive_fp32(N, 1, K, x.data_ptr(), packed.data_ptr(), state.absmax.data_ptr(), state.code.data_ptr(),
                                                  y.data_ptr(), N, (K + 1) // 2, N, 64)
        torch.cuda.synchronize()
        got, want = ours[:, k], y
        same = (got.view(torch.int32) == want.view(torch.int32)) | ((got == 0) & (want == 0))  # the sum of +-0 terms loses the sign of zero
        assert bool(same.all()), f"column {k}: {int((~same).sum())} of {N} values differ from the reference's code[nib] * absmax"
    # 16-bit outputs: the correctly rounded fp32 values
    for dt in (torch.bfloat16, torch.float16):
        st = q.QuantState(absmax=state.absmax, shape=state.shape, code=state.code, blocksize=64, quant_type="nf4", dtype=dt)
        assert torch.equal(q.dequantize_4bit(packed, st).t().contiguous(), ours.to(dt))


@pytest.mark.parametrize("dtype", ["float32", "bfloat16", "float16"])
def test_nf4_quantize_stores_a_nearest_table_entry(q, dtype):
    """NF4 quantize has no reference implementation to compare with ("parity unpinned").  What CAN be pinned: with
    the table pinned by the test above, every stored nibble must be a nearest entry of that table to the normalised value
    x * (1 / absmax) (fp32, the FP4 kernel's normalisation, kernels.cu:436,467), and absmax the block's max |x|."""
    N, K = 256, 4096
    W = weight(N, K, TDT[dtype], seed=4)
    packed, state = q.quantize_4bit(W, quant_type="nf4", compress_statistics=False)
    w = W.float().cpu().numpy().reshape(-1, 64)
    absmax = state.absmax.cpu().numpy()
    assert np.array_equal(absmax, np.abs(w).max(axis=1))
    xn = (w * (np.float32(1.0) / absmax)[:, None]).astype(np.float32).reshape(-1)
    p = packed.cpu().numpy().reshape(-1)
    nib = np.stack([p >> 4, p & 15], axis=1).reshape(-1)
    table = state.code.cpu().numpy().astype(np.float64)
    dist = np.abs(xn.astype(np.float64)[:, None] - table[None, :])
    chosen = dist[np.arange(nib.size), nib]
    assert np.all(chosen <= dist.min(axis=1)), "a stored nibble is not a nearest NF4 table entry"
