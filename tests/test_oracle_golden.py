"""Pin the CPU oracle (oracle/q4_oracle.c) against the golden outputs of the REFERENCE's own CUDA kernels.

tests/golden/*.npz were produced on a B200 by tests/golden/make_golden.py from oracle/_ref (the reference's unmodified
csrc/kernels.cu + csrc/ops.cu compiled for sm_100a).  Everything here is bit-exact, including the GEMV, whose oracle
restates the reference kernel's summation order.  Runs on CPU.
"""
import numpy as np

from conftest import iter_cases


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if got.dtype.kind == "f":
        neq = (bits(got) != bits(want)) & ~(np.isnan(got) & np.isnan(want))
    else:
        neq = got != want
    assert not neq.any(), f"{what}: {int(neq.sum())} of {neq.size} differ, first at {np.argwhere(neq)[:4].ravel()}"


def test_quantize_fp4_golden(oracle, golden):
    g = golden("quantize_fp4")
    ncases = 0
    for k, (dtype, blocksize, n, kind) in iter_cases(g):
        packed, absmax = oracle.quantize_blockwise_4bit(g[k + "_in"], int(blocksize), "fp4")
        assert_bits_equal(absmax, g[k + "_absmax"], f"{k} {dtype} bs={blocksize} n={n} {kind}: absmax")
        assert_bits_equal(packed, g[k + "_packed"], f"{k} {dtype} bs={blocksize} n={n} {kind}: packed")
        ncases += 1
    assert ncases >= 30


def test_blockwise_8bit_golden(oracle, golden):
    g = golden("blockwise_8bit")
    assert_bits_equal(oracle.dynamic_map(), g["code"], "dynamic map")
    for k, (name, blocksize) in iter_cases(g):
        q, absmax = oracle.quantize_blockwise_8bit(g[k + "_in"], int(blocksize), g["code"])
        assert_bits_equal(absmax, g[k + "_absmax"], f"{k} {name} bs={blocksize}: absmax")
        assert_bits_equal(q, g[k + "_q"], f"{k} {name} bs={blocksize}: codes")
        deq = oracle.dequantize_blockwise_8bit(g[k + "_q"], g[k + "_absmax"], int(blocksize), g["code"])
        assert_bits_equal(deq, g[k + "_deq"], f"{k} {name} bs={blocksize}: dequantized")


def test_dequantize_fp4_golden(oracle, golden):
    g = golden("dequantize_fp4")
    for k, (dtype, blocksize, n) in iter_cases(g):
        out = oracle.dequantize_blockwise_4bit(g[k + "_packed"], g[k + "_absmax"], int(n), int(blocksize), "fp4", dtype)
        assert_bits_equal(out, g[k + "_out"], f"{k} {dtype} bs={blocksize} n={n}")


def test_gemv_golden_bit_exact(oracle, golden):
    """The oracle's GEMV follows the reference kernel's per-lane order and shuffle-down tree, so it reproduces the
    reference's fp32, fp16 and bf16 outputs exactly (fp32: with the multiply-add contracted, as nvcc compiles it)."""
    g = golden("gemv")
    for k, (dtype, N, K, code_name) in iter_cases(g):
        out = oracle.gemv_4bit(g[k + "_x"], g[k + "_packed"], g[k + "_absmax"], g[k + "_code"], int(N), int(K), 64, dtype,
                               f32_fused=True)
        assert_bits_equal(out, g[k + "_out"], f"{k} {dtype} {N}x{K} {code_name}")


def test_gemv_truth_close_to_reference(oracle, golden):
    g = golden("gemv")
    for k, (dtype, N, K, code_name) in iter_cases(g):
        truth = oracle.gemv_4bit_f64(g[k + "_x"], g[k + "_packed"], g[k + "_absmax"], g[k + "_code"], int(N), int(K), 64)
        tol = {"float32": 2e-6, "float16": 3e-3, "bfloat16": 3e-2}[dtype]
        assert np.abs(g[k + "_out"] - truth).max() <= tol * np.abs(truth).max()


def test_reference_recipe_golden(oracle, golden):
    """core.py:536-576 + :467-499 step by step: 4-bit pack, offset (torch CUDA mean, taken from the fixture), 8-bit
    double-quant, fused absmax decode, dequantize, GEMV."""
    g = golden("linear_fp4_recipe")
    N, K = (int(v) for v in g["shape"])
    st = oracle.quantize_4bit(g["w"].reshape(N, K), 64, "fp4", offset=float(g["offset"]))
    assert_bits_equal(st["packed"], g["packed"], "packed")
    assert_bits_equal(st["absmax_f32"], g["absmax"], "absmax")
    assert_bits_equal((st["absmax_f32"] - np.float32(g["offset"])).astype(np.float32), g["shifted"], "absmax - offset")
    assert_bits_equal(st["qabsmax"], g["qabsmax"], "qabsmax")
    assert_bits_equal(st["absmax2"], g["absmax2"], "absmax2")
    assert_bits_equal(oracle.state_absmax(st), g["absmax_roundtrip"], "decoded absmax")
    assert_bits_equal(oracle.dequantize_4bit(st, "float16").ravel(), g["wdeq"], "dequantized weight")
    y = oracle.gemv_4bit(g["x"], st["packed"], oracle.state_absmax(st), st["code"], N, K, 64, "float32")
    assert_bits_equal(y, g["y"], "gemv")
    # the numpy float32 mean differs from torch's CUDA reduction only in the last bits
    assert abs(float(st["absmax_f32"].mean(dtype=np.float32)) - float(g["offset"])) < 1e-6
