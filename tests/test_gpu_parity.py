"""GPU parity tests: the CUDA path (through the C ABI / the reference-shaped Python API) against the CPU oracle on the
same seeded inputs, against the golden outputs of the reference's own kernels, and -- at BASELINE.json's full sizes --
through size-independent properties.

Bars (BASELINE.json north_star): quantize (packed bytes, absmax, nested codes) and dequantize outputs BIT-EXACT;
GEMV within max|err| <= 1e-2 * max|truth| of an fp64 truth (stated per test).
"""
import zlib

import numpy as np
import pytest
import torch

from conftest import iter_cases

pytestmark = pytest.mark.gpu

TDT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}
DEV = "cuda:0"


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def f32(t):
    return t.detach().float().cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if got.dtype.kind == "f":
        neq = bits(got) != bits(want)
        both_nan = np.isnan(got) & np.isnan(want)
        neq &= ~both_nan
    else:
        neq = got != want
    assert not neq.any(), f"{what}: {int(neq.sum())} of {neq.size} differ, first at {np.argwhere(neq)[:4].ravel()}"


@pytest.fixture(scope="module")
def q():
    import quantizations_b200 as q

    q._lib.lib()  # fail loudly if the CUDA library is missing
    return q


def make_state(q, st, device=DEV):
    """oracle state dict -> product QuantState on the GPU"""
    code = dev(st["code"])
    if "qabsmax" in st:
        s2 = q.QuantState(absmax=dev(st["absmax2"]), code=dev(st["code2"]), blocksize=256, dtype=torch.float32)
        return q.QuantState(absmax=dev(st["qabsmax"]), shape=torch.Size(st["shape"]), code=code, blocksize=st["blocksize"],
                            quant_type=st["quant_type"], dtype=torch.float16,
                            offset=torch.tensor(float(st["offset"]), dtype=torch.float32, device=device), state2=s2)
    return q.QuantState(absmax=dev(st["absmax_f32"]), shape=torch.Size(st["shape"]), code=code, blocksize=st["blocksize"],
                        quant_type=st["quant_type"], dtype=torch.float16)


# ------------------------------------------------------------------------------------------------ quantize


@pytest.mark.parametrize("dtype", ["float16", "bfloat16", "float32"])
@pytest.mark.parametrize("quant_type", ["fp4", "nf4"])
@pytest.mark.parametrize("blocksize,n", [(64, 64 * 1024), (64, 1001), (64, 63), (64, 1), (128, 5000), (256, 4096),
                                         (512, 5003), (1024, 8192), (2048, 9000), (4096, 16384 + 7)])
def test_quantize_4bit_bit_exact_vs_oracle(q, oracle, dtype, quant_type, blocksize, n):
    rng = np.random.default_rng(zlib.crc32(repr((dtype, quant_type, blocksize, n)).encode()))
    a = (rng.standard_normal(n) * 0.02).astype(np.float32)
    a[rng.integers(0, n, max(1, n // 50))] = 0.0
    A = dev(a, TDT[dtype])
    a = f32(A)
    packed, state = q.quantize_4bit(A, blocksize=blocksize, quant_type=quant_type, compress_statistics=False)
    o_packed, o_absmax = oracle.quantize_blockwise_4bit(a, blocksize, quant_type)
    assert packed.shape == ((n + 1) // 2, 1) and packed.dtype == torch.uint8
    assert_bits_equal(state.absmax.cpu().numpy(), o_absmax, "absmax")
    assert_bits_equal(packed.cpu().numpy().ravel(), o_packed, "packed")


@pytest.mark.parametrize("quant_type", ["fp4", "nf4"])
def test_quantize_4bit_nested_bit_exact_vs_oracle(q, oracle, quant_type):
    """Whole quantize_4bit recipe (core.py:536-576): the offset is torch's CUDA mean, handed to the oracle."""
    rng = np.random.default_rng(7)
    N, K = 512, 1024
    A = dev((rng.standard_normal((N, K)) * 0.02).astype(np.float32), torch.bfloat16)
    packed, state = q.quantize_4bit(A, quant_type=quant_type)
    assert state.nested and state.absmax.dtype == torch.uint8 and state.state2.blocksize == 256
    st = oracle.quantize_4bit(f32(A), 64, quant_type, offset=float(state.offset.item()))
    assert_bits_equal(packed.cpu().numpy().ravel(), st["packed"], "packed")
    assert_bits_equal(state.absmax.cpu().numpy(), st["qabsmax"], "qabsmax")
    assert_bits_equal(state.state2.absmax.cpu().numpy(), st["absmax2"], "absmax2")
    assert_bits_equal(state.state2.code.cpu().numpy(), st["code2"], "code2")
    assert_bits_equal(state.code.cpu().numpy(), st["code"], "code")
    assert tuple(state.shape) == (N, K) and state.dtype == torch.bfloat16


@pytest.mark.parametrize("dtype", ["float16", "bfloat16", "float32"])
@pytest.mark.parametrize("quant_type", ["fp4", "nf4"])
def test_quantize_4bit_binned_encoder_edge_cases(q, oracle, dtype, quant_type):
    """The fast quantize kernel (n % 8 == 0, blocksize <= 256) encodes through a binned table instead of the reference's
    compare tree (csrc/q4_encode_lut.h; the method is swept over all 2^32 floats on the CPU by tests/test_host.py).  Here
    the kernel itself: every 16-bit pattern (or 2^20 random fp32 patterns) as an element, blocks with absmax exactly 1 so that
    the normalised value IS the element, values within a few ulp of every threshold under several absmax scalings, and blocks
    whose absmax is zero / denormal / inf / NaN-polluted (the kernel's compare-tree branch)."""
    rng = np.random.default_rng(zlib.crc32(repr((dtype, quant_type)).encode()))
    tdt = TDT[dtype]
    if dtype == "float32":
        pat = rng.integers(0, 2**32, 1 << 20, dtype=np.uint64).astype(np.uint32).view(np.float32)
    else:
        pat = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(tdt).float().numpy()
    chunks = []
    unit = pat[np.isnan(pat) | (np.abs(pat) <= 1.0)]
    for i in range(0, len(unit), 63):                      # 63 patterns + a 1.0: absmax == 1, x == element
        blk = np.ones(64, np.float32)
        blk[: len(unit[i:i + 63])] = unit[i:i + 63]
        chunks.append(blk)
    chunks.append(np.resize(pat, (len(pat) + 63) // 64 * 64))  # raw patterns: inf / NaN / huge blocks too
    if quant_type == "nf4":  # the 15 midpoints of the table (kernels.cu:851)
        t64 = oracle.nf4_table().astype(np.float64)
        thr = ((t64[:-1] + t64[1:]) / 2).astype(np.float32)
    else:                    # kernels.cu:141-159
        thr = np.array([0.00260417, 0.0859375, 0.20833333, 0.29166667, 0.4166667, 0.583333, 0.8333333], np.float32)
    near = []
    for scale in (1.0, 0.02, 3.0, 1e-3, 7.7):
        for t in np.concatenate([thr, -thr]):
            c = np.float32(t) * np.float32(scale)
            v = c
            for _ in range(4):
                v = np.nextafter(v, np.float32(-np.inf), dtype=np.float32)
            for _ in range(9):
                near.append(v)
                v = np.nextafter(v, np.float32(np.inf), dtype=np.float32)
    near = np.array(near, np.float32)
    for scale, grp in zip((1.0, 0.02, 3.0, 1e-3, 7.7), np.split(near, 5)):
        for i in range(0, len(grp), 63):
            blk = np.full(64, scale, np.float32)
            blk[: len(grp[i:i + 63])] = grp[i:i + 63]
            chunks.append(blk)
    special = np.zeros((6, 64), np.float32)
    special[1, 3] = 1e-41                                   # denormal absmax: 1 / absmax overflows
    special[2, :] = 1e-39
    special[3, 5], special[3, 6] = np.inf, 1.0
    special[4, 7], special[4, 8] = np.nan, -0.5
    special[5, :] = -0.0
    chunks.append(special.ravel())
    a = np.concatenate(chunks).astype(np.float32)
    A = dev(a, tdt)
    a = f32(A)
    assert a.size % 8 == 0
    for blocksize in (64, 256):
        n = a.size // blocksize * blocksize
        packed, state = q.quantize_4bit(A[:n], blocksize=blocksize, quant_type=quant_type, compress_statistics=False)
        o_packed, o_absmax = oracle.quantize_blockwise_4bit(a[:n], blocksize, quant_type)
        assert_bits_equal(state.absmax.cpu().numpy(), o_absmax, f"absmax bs{blocksize}")
        assert_bits_equal(packed.cpu().numpy().ravel(), o_packed, f"packed bs{blocksize}")


def test_quantize_fp4_matches_reference_golden(q, golden):
    g = golden("quantize_fp4")
    for k, (dtype, blocksize, n, kind) in iter_cases(g):
        A = dev(g[k + "_in"], TDT[dtype])
        packed, state = q.quantize_4bit(A, blocksize=int(blocksize), quant_type="fp4", compress_statistics=False)
        assert_bits_equal(state.absmax.cpu().numpy(), g[k + "_absmax"], f"{k} {dtype} bs={blocksize} n={n} {kind}: absmax")
        assert_bits_equal(packed.cpu().numpy().ravel(), g[k + "_packed"], f"{k} {dtype} bs={blocksize} n={n} {kind}: packed")


def test_blockwise_8bit_matches_reference_golden(q, golden):
    g = golden("blockwise_8bit")
    for k, (name, blocksize) in iter_cases(g):
        A = dev(g[k + "_in"])
        out, st = q.quantize_blockwise(A, blocksize=int(blocksize))
        assert_bits_equal(st.absmax.cpu().numpy(), g[k + "_absmax"], f"{k} {name} bs={blocksize}: absmax")
        assert_bits_equal(out.cpu().numpy(), g[k + "_q"], f"{k} {name} bs={blocksize}: codes")
        deq = q.dequantize_blockwise(out, st)
        assert_bits_equal(deq.cpu().numpy(), g[k + "_deq"], f"{k} {name} bs={blocksize}: dequantized")


@pytest.mark.parametrize("blocksize,n", [(256, 256 * 100 + 3), (4096, 4096 * 3 + 1), (64, 1000), (1024, 1024)])
def test_blockwise_8bit_bit_exact_vs_oracle(q, oracle, blocksize, n):
    rng = np.random.default_rng(blocksize + n)
    a = rng.uniform(-1, 1, n).astype(np.float32) * 0.3
    out, st = q.quantize_blockwise(dev(a), blocksize=blocksize)
    o_q, o_am = oracle.quantize_blockwise_8bit(a, blocksize)
    assert_bits_equal(st.absmax.cpu().numpy(), o_am, "absmax")
    assert_bits_equal(out.cpu().numpy(), o_q, "codes")
    deq = q.dequantize_blockwise(out, st)
    assert_bits_equal(deq.cpu().numpy(), oracle.dequantize_blockwise_8bit(o_q, o_am, blocksize), "dequantized")


# ------------------------------------------------------------------------------------------------ dequantize


@pytest.mark.parametrize("dtype", ["float16", "bfloat16", "float32"])
@pytest.mark.parametrize("quant_type", ["fp4", "nf4"])
@pytest.mark.parametrize("nested", [False, True])
@pytest.mark.parametrize("shape", [(256, 512), (3, 100), (1, 7), (64, 64)])
def test_dequantize_4bit_bit_exact_vs_oracle(q, oracle, dtype, quant_type, nested, shape):
    rng = np.random.default_rng(11)
    w = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    A = dev(w, TDT[dtype])
    packed, state = q.quantize_4bit(A, quant_type=quant_type, compress_statistics=nested)
    out = q.dequantize_4bit(packed, state)
    assert out.shape == (shape[1], shape[0]) and out.dtype == TDT[dtype]  # the reference returns out.t() (core.py:634)
    st = oracle.quantize_4bit(f32(A), 64, quant_type, offset=float(state.offset.item()) if nested else None,
                              compress_statistics=nested)
    want = oracle.dequantize_4bit(st, dtype)
    assert_bits_equal(f32(out.t()), want, "dequantized weight")


def test_dequantize_fp4_matches_reference_golden(q, golden):
    g = golden("dequantize_fp4")
    for k, (dtype, blocksize, n) in iter_cases(g):
        n, blocksize = int(n), int(blocksize)
        state = q.QuantState(absmax=dev(g[k + "_absmax"]), shape=torch.Size((1, n)), code=q.get_4bit_type("fp4", DEV),
                             blocksize=blocksize, quant_type="fp4", dtype=TDT[dtype])
        out = q.dequantize_4bit(dev(g[k + "_packed"]), state, blocksize=blocksize)
        assert_bits_equal(f32(out).ravel(), g[k + "_out"], f"{k} {dtype} bs={blocksize} n={n}")


def test_reference_recipe_golden(q, golden):
    """The reference's full quantize_4bit -> dequantize_4bit / gemv_4bit recipe on one weight, step by step."""
    g = golden("linear_fp4_recipe")
    N, K = (int(v) for v in g["shape"])
    W = dev(g["w"], torch.float16).reshape(N, K)
    packed, state = q.quantize_4bit(W, quant_type="fp4")
    assert_bits_equal(packed.cpu().numpy().ravel(), g["packed"], "packed")
    assert_bits_equal(np.float32(state.offset.item()), g["offset"], "offset (torch CUDA mean)")
    assert_bits_equal(state.absmax.cpu().numpy(), g["qabsmax"], "qabsmax")
    assert_bits_equal(state.state2.absmax.cpu().numpy(), g["absmax2"], "absmax2")
    wdeq = q.dequantize_4bit(packed, state).t()
    assert_bits_equal(f32(wdeq).ravel(), g["wdeq"], "dequantized weight (fused double-quant decode)")
    x = dev(g["x"]).reshape(1, 1, K)
    y = q.gemv_4bit(x, packed, state=state)
    assert y.shape == (1, 1, N)
    np.testing.assert_allclose(f32(y).ravel(), g["y"], rtol=0, atol=2e-6 * np.abs(g["y"]).max() + 1e-7)


# ------------------------------------------------------------------------------------------------ GEMV


def gemv_truth(oracle, x, st, N, K):
    return oracle.gemv_4bit_f64(x, st["packed"], oracle.state_absmax(st), st["code"], N, K, st["blocksize"])


@pytest.mark.parametrize("dtype", ["bfloat16", "float16", "float32"])
@pytest.mark.parametrize("quant_type", ["nf4", "fp4"])
@pytest.mark.parametrize("nested", [True, False])
@pytest.mark.parametrize("N,K", [(64, 256), (33, 1088), (1024, 4096), (7, 14336), (300, 2048), (5, 64), (129, 6144)])
def test_gemv_vs_fp64_truth(q, oracle, dtype, quant_type, nested, N, K):
    """tolerance: max|y - truth| <= 1e-2 * max|truth| (BASELINE north_star), plus the output type's own rounding."""
    rng = np.random.default_rng(N * 131 + K)
    W = dev((rng.standard_normal((N, K)) * 0.02).astype(np.float32), TDT[dtype] if dtype != "float32" else torch.float16)
    packed, state = q.quantize_4bit(W, quant_type=quant_type, compress_statistics=nested)
    x = dev(rng.standard_normal(K).astype(np.float32), TDT[dtype]).reshape(1, 1, K)
    bias = dev(rng.standard_normal(N).astype(np.float32), TDT[dtype])
    y = q.gemv_4bit(x, packed, state=state)
    yb = q.gemv_4bit(x, packed, state=state, bias=bias)
    assert y.shape == (1, 1, N) and y.dtype == TDT[dtype]
    st = oracle.quantize_4bit(f32(W), 64, quant_type, offset=float(state.offset.item()) if nested else None,
                              compress_statistics=nested)
    truth = gemv_truth(oracle, f32(x).ravel(), st, N, K)
    scale = np.abs(truth).max()
    err = np.abs(f32(y).ravel() - truth).max()
    assert err <= 1e-2 * scale, f"max err {err:.3e} vs scale {scale:.3e}"
    tight = {"float32": 1e-5, "float16": 2e-3, "bfloat16": 8e-3}[dtype]
    assert err <= tight * scale, f"max err {err:.3e} exceeds the expected {tight} * {scale:.3e}"
    errb = np.abs(f32(yb).ravel() - (truth + f32(bias))).max()
    assert errb <= 1e-2 * np.abs(truth + f32(bias)).max()


def test_gemv_matches_reference_golden(q, golden):
    """Against the reference kernel's own outputs (fp32 exported instance; fp16/bf16 instances via the test shim).
    The reference's 16-bit instances round code, absmax, code*absmax and every product to T (kernels.cu:1120,1131,
    1169,1206); ours keeps fp32 where it can, so grade ours against the reference with the reference's own distance
    from truth as the yard-stick."""
    g = golden("gemv")
    for k, (dtype, N, K, code_name) in iter_cases(g):
        N, K = int(N), int(K)
        state = q.QuantState(absmax=dev(g[k + "_absmax"]), shape=torch.Size((N, K)), code=dev(g[k + "_code"]), blocksize=64,
                             quant_type=code_name, dtype=torch.float16)
        x = dev(g[k + "_x"], TDT[dtype]).reshape(1, 1, K)
        y = f32(q.gemv_4bit(x, dev(g[k + "_packed"]), state=state)).ravel()
        ref = g[k + "_out"]
        scale = np.abs(ref).max()
        tol = {"float32": 2e-6, "float16": 4e-3, "bfloat16": 3e-2}[dtype]
        assert np.abs(y - ref).max() <= tol * scale, f"{k} {dtype} {N}x{K} {code_name}: {np.abs(y - ref).max():.3e} vs {scale:.3e}"


def test_gemv_error_behaviour(q):
    """reference core.py:453-460"""
    W = torch.randn(64, 128, device=DEV, dtype=torch.float16)
    packed, state = q.quantize_4bit(W)
    with pytest.raises(ValueError):
        q.gemv_4bit(torch.randn(1, 1, 128, device=DEV), packed, state=None)
    with pytest.raises(ValueError):
        q.gemv_4bit(torch.randn(2, 128, device=DEV), packed, state=state)


# ------------------------------------------------------------------------------------------------ full-size properties


LLAMA3_8B_SHAPES = [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336)]


@pytest.mark.parametrize("N,K", LLAMA3_8B_SHAPES)
@pytest.mark.parametrize("quant_type", ["nf4", "fp4"])
def test_full_size_roundtrip_and_linearity(q, N, K, quant_type):
    """BASELINE.json config 2 sizes.  Properties that need no CPU oracle:
    (1) quantize is idempotent on its own output: quantize(dequantize(q(W))) reproduces the same packed bytes;
    (2) every dequantized block's max |value| equals its decoded absmax times a code of magnitude 1 (the absmax
        element always maps to +-1);
    (3) GEMV is linear: gemv(a*x1 + x2) == a*gemv(x1) + gemv(x2) within fp rounding, and
    (4) GEMV agrees with a torch fp32 matmul of the dequantized weight (max err <= 1e-2 * max|y|)."""
    torch.manual_seed(0)
    W = (torch.randn(N, K, device=DEV, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    packed, state = q.quantize_4bit(W, quant_type=quant_type)
    assert packed.numel() == N * K // 2 and state.absmax.numel() == N * K // 64
    Wd = q.dequantize_4bit(packed, state).t().contiguous()
    assert Wd.shape == (N, K) and Wd.dtype == torch.bfloat16
    # (2)
    blockmax = Wd.float().reshape(-1, 64).abs().amax(dim=1)
    absmax = q.dequantize_blockwise(state.absmax, state.state2) + state.offset
    assert torch.equal(blockmax.to(torch.bfloat16), absmax.abs().to(torch.bfloat16))
    # (1) re-quantising the dequantised weight with UNCOMPRESSED statistics must give back the same codes wherever the
    # bf16 rounding of absmax did not move the block maximum (it is the same value by (2)), i.e. everywhere.
    packed2, state2 = q.quantize_4bit(Wd, quant_type=quant_type, compress_statistics=False)
    p1, p2 = packed, packed2
    if quant_type == "fp4":
        # FP4 code 0b1000 decodes to -0.0, whose sign the quantiser does not see (x < 0 is false): it re-encodes as 0b0000
        def fold(p):
            hi, lo = p >> 4, p & 0xF
            return (torch.where(hi == 8, torch.zeros_like(hi), hi) << 4) | torch.where(lo == 8, torch.zeros_like(lo), lo)

        p1, p2 = fold(packed), fold(packed2)
    same = (p2 == p1).float().mean().item()
    assert same > 0.999, f"only {same:.5f} of the packed bytes survive a dequantize->quantize round trip"
    # (3) + (4)
    torch.manual_seed(1)
    x1 = torch.randn(1, 1, K, device=DEV, dtype=torch.bfloat16)
    x2 = torch.randn(1, 1, K, device=DEV, dtype=torch.bfloat16)
    y1, y2 = q.gemv_4bit(x1, packed, state=state).float(), q.gemv_4bit(x2, packed, state=state).float()
    y12 = q.gemv_4bit((2 * x1 + x2), packed, state=state).float()
    ref = (2 * x1 + x2).float().reshape(1, K) @ Wd.float().t()
    scale = ref.abs().max().item()
    assert (y12 - ref).abs().max().item() <= 1e-2 * scale
    assert (y12 - (2 * y1 + y2)).abs().max().item() <= 2e-2 * scale


def test_linear4bit_module_decode_and_prefill(q):
    """Linear4bit as HF constructs it (modules.py:86-110): quantises on .to('cuda'), decode -> GEMV, prefill -> GEMM."""
    torch.manual_seed(0)
    lin = q.Linear4bit(512, 256, bias=True, compute_dtype=torch.bfloat16, compress_statistics=True, quant_type="nf4")
    w_ref = lin.weight.data.clone()
    lin = lin.to(DEV)
    assert type(lin.weight).__name__ == "Params4bit" and lin.weight.dtype == torch.uint8
    assert lin.weight.shape == (512 * 256 // 2, 1) and lin.quant_state is lin.weight.quant_state
    Wd = q.dequantize_4bit(lin.weight.data, lin.weight.quant_state).t().float()
    assert (Wd.cpu() - w_ref).abs().max() < 0.2 * w_ref.abs().max()
    x = torch.randn(1, 1, 512, device=DEV, dtype=torch.bfloat16)
    y = lin(x)
    ref = x.float() @ Wd.t() + lin.bias.float()
    assert y.shape == (1, 1, 256) and y.dtype == torch.bfloat16
    assert (y.float() - ref).abs().max() <= 1e-2 * ref.abs().max()
    xp = torch.randn(2, 17, 512, device=DEV, dtype=torch.bfloat16)
    yp = lin(xp)
    refp = xp.float() @ Wd.t() + lin.bias.float()
    assert yp.shape == (2, 17, 256)
    assert (yp.float() - refp).abs().max() <= 2e-2 * refp.abs().max()
    # the accelerate/HF rebuild pattern: Params4bit(value, requires_grad=False, **old.__dict__).to(device)
    old = lin.weight
    rebuilt = q.Params4bit(old.data, requires_grad=False, **old.__dict__).to(DEV)
    assert rebuilt.quant_state is old.quant_state and rebuilt.bnb_quantized


# ------------------------------------------------------------------------------------------------ prefill (fused tcgen05 GEMM)


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("quant_type,nested", [("fp4", True), ("nf4", True), ("nf4", False)])
@pytest.mark.parametrize("M,N,K", [(16, 128, 64), (1, 256, 256), (33, 384, 512), (100, 200, 1024), (128, 1024, 4096),
                                   (300, 512, 2048), (17, 4096, 4096), (64, 1024, 14336), (700, 640, 1024), (1000, 1024, 2048)])
def test_fused_gemm_vs_fp64_truth(q, oracle, dtype, quant_type, nested, M, N, K):
    """gemm_4bit (dequantise fused into tcgen05 MMA) against an fp64 product of the ORACLE's dequantised weight.
    tolerance: max|y - truth| <= 1e-2 * max|truth|; ragged M and N (not multiples of the 16..256 x 128 tiles) included."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    W = dev((rng.standard_normal((N, K)) * 0.02).astype(np.float32), TDT[dtype])
    packed, state = q.quantize_4bit(W, quant_type=quant_type, compress_statistics=nested)
    X = dev(rng.standard_normal((M, K)).astype(np.float32), TDT[dtype])
    bias = dev(rng.standard_normal(N).astype(np.float32), TDT[dtype])
    y = q.gemm_4bit(X.reshape(1, M, K), packed, state, bias=bias)
    assert y.shape == (1, M, N) and y.dtype == TDT[dtype]
    st = oracle.quantize_4bit(f32(W), 64, quant_type, offset=float(state.offset.item()) if nested else None,
                              compress_statistics=nested)
    wdeq = oracle.dequantize_4bit(st, "float32")
    truth = f32(X).astype(np.float64) @ wdeq.astype(np.float64).T + f32(bias).astype(np.float64)
    err = np.abs(f32(y).reshape(M, N) - truth).max()
    assert err <= 1e-2 * np.abs(truth).max(), f"max err {err:.3e} vs scale {np.abs(truth).max():.3e}"
    y0 = q.gemm_4bit(X, packed, state)
    np.testing.assert_allclose(f32(y0), truth - f32(bias), rtol=0, atol=1.2e-2 * np.abs(truth).max())


def test_fused_gemm_matches_gemv_and_cublas_path(q, monkeypatch):
    """The three routes through matmul_4bit agree: GEMV (one token), fused GEMM, dequantise + F.linear."""
    torch.manual_seed(3)
    lin = q.Linear4bit(1024, 768, bias=True, compute_dtype=torch.float16, quant_type="nf4").to(DEV)
    lin.bias.data = lin.bias.data.half()
    x = torch.randn(1, 24, 1024, device=DEV, dtype=torch.float16)
    monkeypatch.setenv("Q4_PREFILL", "fused")
    y_fused = lin(x)
    monkeypatch.setenv("Q4_PREFILL", "cublas")
    y_cublas = lin(x)
    y_gemv = torch.cat([lin(x[:, i : i + 1]) for i in range(24)], dim=1)
    scale = y_cublas.float().abs().max().item()
    assert (y_fused.float() - y_cublas.float()).abs().max().item() <= 5e-3 * scale
    assert (y_fused.float() - y_gemv.float()).abs().max().item() <= 5e-3 * scale


def test_fused_gemm_error_behaviour(q):
    W = torch.randn(128, 192, device=DEV, dtype=torch.float16)
    packed, state = q.quantize_4bit(W)
    with pytest.raises(ValueError):
        q.gemm_4bit(torch.randn(4, 128, device=DEV, dtype=torch.float16), packed, state)
    with pytest.raises(NotImplementedError):
        q.gemm_4bit(torch.randn(4, 192, device=DEV, dtype=torch.float32), packed, state)
    packed128, state128 = q.quantize_4bit(W, blocksize=128)
    with pytest.raises(NotImplementedError):
        q.gemm_4bit(torch.randn(4, 192, device=DEV, dtype=torch.float16), packed128, state128)


# ------------------------------------------------------------------------------------------------ end-to-end decode loop


def test_llama_decode_matches_dense_model_with_dequantised_weights(q):
    """The whole chain inside a real decoder (quantizations_b200/llama.py): prefill through the fused GEMM / cuBLAS path,
    decode through the GEMV, eager and CUDA-graph replay -- against the same model with dense nn.Linear layers holding the
    DEQUANTISED weights (so the only differences are kernel arithmetic, not quantisation error)."""
    from quantizations_b200 import llama

    cfg = llama.LlamaConfig(hidden=512, inter=1024, layers=2, heads=8, kv_heads=2, vocab=1000, max_len=64)
    dev_ = torch.device(DEV)
    mq = llama.Llama(cfg, llama.linear4bit_factory(dev_, torch.bfloat16, "nf4"), dev_, torch.bfloat16)

    def dense_from(mod):
        W = q.dequantize_4bit(mod.weight.data, mod.weight.quant_state).t().contiguous()
        lin = torch.nn.Linear(W.shape[1], W.shape[0], bias=False, device=dev_, dtype=torch.bfloat16)
        lin.weight.data = W
        return lin

    md = llama.Llama(cfg, llama.dense_factory(dev_, torch.bfloat16), dev_, torch.bfloat16)
    for Lq, Ld in zip(mq.layers, md.layers):
        for name in ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"):
            setattr(Ld, name, dense_from(getattr(Lq, name)))
    md.embed, md.lm_head = mq.embed, mq.lm_head
    prompt = torch.arange(1, 20, device=dev_)
    lq = mq.forward(prompt, torch.arange(19, device=dev_)).float()
    ld = md.forward(prompt, torch.arange(19, device=dev_)).float()
    assert (lq - ld).abs().max().item() <= 3e-2 * ld.abs().max().item()
    # one decode step on top of the prefilled caches
    tok, pos = torch.tensor([5], device=dev_), torch.tensor([19], device=dev_)
    sq, sd = mq.forward(tok, pos).float(), md.forward(tok, pos).float()
    assert (sq - sd).abs().max().item() <= 3e-2 * sd.abs().max().item()
    # greedy generation: graph replay reproduces the eager tokens exactly (same kernels, same order)
    t_eager, _ = mq.generate(prompt, 12, use_graph=False)
    t_graph, _ = mq.generate(prompt, 12, use_graph=True)
    assert torch.equal(t_eager, t_graph)


@pytest.mark.parametrize("nested", [True, False])
@pytest.mark.parametrize("quant_type", ["nf4", "fp4"])
def test_grouped_gemv_equals_member_gemvs(q, nested, quant_type):
    """Linear4bitGroup (q/k/v in one launch) returns bit-identical outputs to the members called one by one: same kernel
    arithmetic per row, only the launch is shared; the members keep working on their re-laid (view) storage."""
    torch.manual_seed(5)
    K, Ns = 1024, [1024, 256, 512]
    lins = []
    for n in Ns:
        lin = q.Linear4bit(K, n, bias=False, compute_dtype=torch.bfloat16, compress_statistics=nested, quant_type=quant_type)
        lins.append(lin.to(DEV))
    x = torch.randn(1, 1, K, device=DEV, dtype=torch.bfloat16)
    before = [lin(x).clone() for lin in lins]
    grp = q.Linear4bitGroup(lins)
    outs = grp(x)
    after = [lin(x) for lin in lins]
    assert [o.shape[-1] for o in outs] == Ns
    for b, o, a_ in zip(before, outs, after):
        assert torch.equal(b, a_), "re-laying the storage changed a member's output"
        assert torch.equal(b, o), "grouped launch differs from the member launch"
    # prefill falls back to the members
    xp = torch.randn(1, 5, K, device=DEV, dtype=torch.bfloat16)
    op = grp(xp)
    assert [o.shape for o in op] == [(1, 5, n) for n in Ns]
    with pytest.raises(ValueError):
        q.Linear4bitGroup([lins[0], q.Linear4bit(512, 64, quant_type=quant_type, compress_statistics=nested).to(DEV)])


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_fused_decode_glue_matches_separate_torch_ops(q, dtype):
    """gemv_4bit_fused: RMSNorm / SwiGLU folded into the activation staging and the residual add into the epilogue give the
    same result as the separate torch kernels followed by the plain GEMV (tolerance: one rounding of the output type)."""
    torch.manual_seed(9)
    dt = TDT[dtype]
    K, N = 4096, 1024
    W = (torch.randn(N, K, device=DEV) * 0.02).to(dt)
    packed, state = q.quantize_4bit(W, quant_type="nf4")
    x = torch.randn(1, 1, K, device=DEV, dtype=dt) * 3
    gamma = (1 + 0.1 * torch.randn(K, device=DEV)).to(dt)
    res = torch.randn(1, 1, N, device=DEV, dtype=dt)
    gate = torch.randn(1, 1, K, device=DEV, dtype=dt)
    tol = {"bfloat16": 1.6e-2, "float16": 2e-3}[dtype]

    ref = q.gemv_4bit(torch.nn.functional.rms_norm(x, (K,), gamma, 1e-5), packed, state=state).float() + res.float()
    got = q.gemv_4bit_fused(x, packed, state, rms_weight=gamma, rms_eps=1e-5, residual=res).float()
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()

    ref = q.gemv_4bit(torch.nn.functional.silu(gate) * x, packed, state=state).float()
    got = q.gemv_4bit_fused(x, packed, state, gate=gate).float()
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()

    # in-place residual stream: out aliases residual
    stream = res.clone()
    q.gemv_4bit_fused(x, packed, state, residual=stream, out=stream)
    ref = q.gemv_4bit(x, packed, state=state).float() + res.float()
    assert (stream.float() - ref).abs().max().item() <= tol * ref.abs().max().item()
    # plain call through the fused entry point is bit-identical to gemv_4bit
    assert torch.equal(q.gemv_4bit_fused(x, packed, state), q.gemv_4bit(x, packed, state=state))


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
@pytest.mark.parametrize("nested", [True, False])
@pytest.mark.parametrize("N,K", [(14336, 4096), (1792, 4096), (512, 8192), (8, 4096)])
def test_swiglu_group_equals_grouped_launch_then_swiglu_staging(q, dtype, nested, N, K):
    """Linear4bitGroup(swiglu=True).forward_swiglu: gate / up rows interleaved in chunks of four, silu(gate) * up formed in the
    launch's epilogue.  Same rounding steps as the grouped gate/up launch followed by the SwiGLU activation staging of the down
    projection, so: bit-identical to silu-staging's input, and the down projection gives bit-identical outputs either way."""
    torch.manual_seed(N + K)
    dt = TDT[dtype]
    lins = []
    for i in range(2):
        lin = q.Linear4bit(K, N, bias=False, compute_dtype=dt, quant_type="nf4", compress_statistics=nested)
        lin.weight = q.Params4bit((torch.randn(N, K) * 0.02).to(dt), requires_grad=False, quant_type="nf4", module=lin,
                                  compress_statistics=nested)
        lins.append(lin.to(DEV))
    down = q.Linear4bit(N, 256, bias=False, compute_dtype=dt, quant_type="nf4")
    down.weight = q.Params4bit((torch.randn(256, N) * 0.02).to(dt), requires_grad=False, quant_type="nf4", module=down)
    down = down.to(DEV)
    x = torch.randn(1, 1, K, device=DEV, dtype=dt)
    gamma = (1 + 0.1 * torch.randn(K, device=DEV)).to(dt)
    sw = q.Linear4bitGroup(lins, swiglu=True)
    g, u = lins[0](x), lins[1](x)
    h = sw.forward_swiglu(x)
    assert h.shape == (1, 1, N)
    want = (torch.nn.functional.silu(g.float()).to(dt).float() * u.float()).to(dt)
    assert (h.float() - want.float()).abs().max().item() <= 2e-2 * want.float().abs().max().item()
    if N % 64 == 0:  # the down projection: plain staging of h == SwiGLU staging of (g, u), bit for bit
        st = down.weight.quant_state
        y_sw = q.gemv_4bit_fused(h, down.weight.data, st)
        y_ref = q.gemv_4bit_fused(u, down.weight.data, st, gate=g)
        assert torch.equal(y_sw, y_ref)
    # with the RMSNorm fused in, against the unfused sequence
    hn = sw.forward_swiglu(x, rms_weight=gamma, rms_eps=1e-5)
    xn = torch.nn.functional.rms_norm(x, (K,), gamma, 1e-5)
    wantn = (torch.nn.functional.silu(lins[0](xn).float()).to(dt).float() * lins[1](xn).float()).to(dt)
    assert (hn.float() - wantn.float()).abs().max().item() <= 3e-2 * wantn.float().abs().max().item()
    with pytest.raises(ValueError):
        sw.forward_fused(x)
    assert all(torch.equal(a, b) for a, b in zip(sw(x), (g, u)))


@pytest.mark.parametrize("next_shape", [(4096, 4096), (28672, 4096), (4096, 14336), (1024, 1792), (8, 512), (300, 2048)])
def test_prefetch_hint_never_changes_the_result(q, next_shape):
    """q4_gemv_fused_t.prefetch / prefetch_K is a hint: whatever the next weight looks like (more or fewer row tiles than CTAs,
    ragged k tiles -> the hint is dropped, a bare tensor -> the whole-range form), the output is bit-identical."""
    torch.manual_seed(11)
    N, K = 4096, 4096
    W = (torch.randn(N, K, device=DEV) * 0.02).to(torch.bfloat16)
    packed, state = q.quantize_4bit(W, quant_type="nf4")
    x = torch.randn(1, 1, K, device=DEV, dtype=torch.bfloat16)
    nN, nK = next_shape
    nxt = torch.randint(0, 255, (nN * nK // 2, 1), device=DEV, dtype=torch.uint8)
    want = q.gemv_4bit_fused(x, packed, state)
    for hint in ((nxt, nK), nxt, (nxt, 0)):
        got = q.gemv_4bit_fused(x, packed, state, prefetch=hint)
        assert torch.equal(got, want)
        assert torch.equal(q.gemv_4bit(x, packed, state=state, flags=q._lib.Q4_GEMV_PDL, prefetch=hint), want)
    torch.cuda.synchronize()


def test_linear4bit_cached_descriptor_follows_the_weight(q):
    """Linear4bit.forward keeps a cached launch descriptor for single-vector calls; it must notice a new weight, a changed launch
    flag, a prefetch hint and another activation dtype, and must agree with core.gemv_4bit every time."""
    torch.manual_seed(12)
    K, N = 1024, 512
    lin = q.Linear4bit(K, N, bias=True, compute_dtype=torch.bfloat16, quant_type="nf4")
    W1 = (torch.randn(N, K) * 0.02).to(torch.bfloat16)
    lin.weight = q.Params4bit(W1, requires_grad=False, quant_type="nf4", module=lin)
    lin.bias.data = torch.randn(N).to(torch.bfloat16)
    lin = lin.to(DEV)
    x = torch.randn(1, 1, K, device=DEV, dtype=torch.bfloat16)

    def direct():
        return q.gemv_4bit(x, lin.weight.data, state=lin.weight.quant_state, bias=lin.bias, flags=lin.gemv_flags)

    y1 = lin(x)
    assert torch.equal(y1, direct()) and torch.equal(lin(x), y1)
    lin.gemv_flags = 0
    assert torch.equal(lin(x), direct())
    lin.prefetch_next = (lin.weight.data, K)
    assert torch.equal(lin(x), y1)
    W2 = (torch.randn(N, K) * 0.02).to(torch.bfloat16)
    lin.weight = q.Params4bit(W2, requires_grad=False, quant_type="fp4", module=lin).to(DEV)
    y2 = lin(x)
    assert torch.equal(y2, direct()) and not torch.equal(y2, y1)
    xh = x.to(torch.float16)   # fp16 input keeps the configured bf16 compute dtype (reference modules.py:112-122): cast in, cast out
    yh = lin(xh)
    assert yh.dtype == torch.float16 and torch.equal(yh, lin(xh.to(torch.bfloat16)).to(torch.float16))


@pytest.mark.parametrize("shape", [(4096, 4096), (14336, 4096), (4096, 14336), (1000, 512), (136, 256)])
@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_tcgen05_gemv_matches_mma_sync_gemv(q, shape, dtype):
    """The tcgen05 decode kernel (q4_gemv_tc.cuh: lut + workspace given) and the mma.sync kernel (no workspace) compute the
    same sums in a different order: outputs agree to the rounding of the output type, on repeated calls with one workspace
    (its counters must return to zero), including shapes whose last row tile is ragged and after a launch of another shape."""
    import ctypes

    from quantizations_b200 import _lib

    N, K = shape
    dt = TDT[dtype]
    torch.manual_seed(3)
    W = (torch.randn(N, K, device=DEV) * 0.02).to(dt)
    packed, st = q.quantize_4bit(W, quant_type="nf4")
    stats, lut = st.native_stats(), st.lut(dt)
    ws = torch.zeros(_lib.Q4_GEMV_WORKSPACE_BYTES, dtype=torch.uint8, device=DEV)
    ws[65536:].fill_(0x7F)  # the partial-sum area may hold anything; only the counters must start at zero
    bias = torch.randn(N, device=DEV, dtype=dt)
    for rep in range(3):
        x = torch.randn(1, 1, K, device=DEV, dtype=dt)
        outs = []
        for w in (ws, None):
            out = torch.full((N,), 3.0, device=DEV, dtype=dt)
            f = _lib.GemvFused(x.data_ptr(), None, None, 0.0, packed.data_ptr(), ctypes.pointer(stats), None, None, 1, st.code.data_ptr(),
                               bias.data_ptr(), out.data_ptr(), N, K, 64, {"bfloat16": _lib.Q4_BF16, "float16": _lib.Q4_F16}[dtype], 0, None, 0,
                               lut.data_ptr(), None if w is None else w.data_ptr(), 0 if w is None else w.numel())
            assert _lib.lib().q4_gemv_4bit_fused(ctypes.byref(f), torch.cuda.current_stream().cuda_stream) == 0
            outs.append(out.float())
        tol = {"bfloat16": 1e-2, "float16": 2e-3}[dtype] * outs[1].abs().max().item()
        assert (outs[0] - outs[1]).abs().max().item() <= tol
        assert int(ws[:65536].view(torch.int32).ne(0).sum()) == 0


@pytest.mark.parametrize("pos", [0, 1, 7, 64, 65, 200, 255])
def test_decode_attention_early_cache_reads_change_nothing(q, pos):
    """Q4_ATTN_EARLY_CACHE only moves loads of data written by earlier steps in front of griddepcontrol.wait and fetches the
    cached rows in batches of 8 per warp: the output and the appended cache row must be bit-identical to the plain launch, at
    context lengths that leave batches empty, partly filled and repeated."""
    from quantizations_b200 import _lib
    from quantizations_b200.core import decode_attention

    nh, nkv, hd, max_len = 8, 2, 128, 256
    g = torch.Generator(device=DEV).manual_seed(pos)
    dt = torch.bfloat16
    qkv = torch.randn(1, 1, (nh + 2 * nkv) * hd, device=DEV, generator=g).to(dt)
    ang = torch.rand(max_len, hd // 2, device=DEV, generator=g) * 6.0
    cos, sin = ang.cos().to(dt), ang.sin().to(dt)
    kc = torch.randn(nkv, max_len, hd, device=DEV, generator=g).to(dt)
    vc = torch.randn(nkv, max_len, hd, device=DEV, generator=g).to(dt)
    p = torch.tensor([pos], device=DEV)
    outs = []
    for flags in (_lib.Q4_GEMV_PDL, _lib.Q4_GEMV_PDL | _lib.Q4_ATTN_EARLY_CACHE, 0):
        k, v = kc.clone(), vc.clone()
        o = decode_attention(qkv, cos, sin, k, v, p, nh, nkv, flags=flags)
        outs.append((o.clone(), k, v))
    for o, k, v in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(k, outs[0][1]) and torch.equal(v, outs[0][2])
    # and against plain torch attention over positions [0, pos] (fp32 softmax), within bf16 rounding of the output
    q_ = qkv.view(-1)[: nh * hd].view(nh, hd).float()
    kn = outs[0][1][:, : pos + 1].float()
    vn = outs[0][2][:, : pos + 1].float()
    c, s_ = cos[pos].float(), sin[pos].float()
    q1, q2 = q_[:, : hd // 2], q_[:, hd // 2:]
    rnd = lambda t: t.to(dt).float()  # noqa: E731
    qr = torch.cat((rnd(rnd(q1 * c) - rnd(q2 * s_)), rnd(rnd(q2 * c) + rnd(q1 * s_))), dim=-1)
    want = torch.empty(nh, hd, device=DEV)
    for h in range(nh):
        sc = (kn[h // (nh // nkv)] @ qr[h]) / hd**0.5
        want[h] = torch.softmax(sc, dim=0) @ vn[h // (nh // nkv)]
    got = outs[0][0].float().view(nh, hd)
    assert (got - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("dtype", ["bfloat16", "float16", "float32"])
@pytest.mark.parametrize("n", [1, 7, 8, 1000, 128256, 300001])
def test_argmax_matches_torch(q, dtype, n):
    """q4_argmax (greedy sampling glue): the index torch.argmax returns, lowest index on ties, replayable in a CUDA graph."""
    from quantizations_b200.core import argmax

    g = torch.Generator(device=DEV).manual_seed(n)
    x = torch.randn(n, device=DEV, generator=g).to(TDT[dtype])
    assert int(argmax(x)) == int(x.argmax())
    if n > 8:
        x[n // 3] = x[n - 2] = x.max() + 1  # a tie: the lowest index wins
        assert int(argmax(x)) == n // 3
    out = torch.zeros(1, dtype=torch.int64, device=DEV)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        argmax(x, out=out)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            argmax(x, out=out)
    for i in range(3):
        x[(i * 7919) % n] = 1e4 * (i + 1)
        gr.replay()
        torch.cuda.synchronize()
        assert int(out) == (i * 7919) % n


def test_fused_decode_attention_matches_torch_attention_path(q):
    """q4_decode_attention (RoPE + KV append + GQA attention in one launch) against the torch ops it replaces inside the same
    decoder (llama.py, head_dim 128): same logits within bf16 rounding, same cache contents, over several decode steps."""
    from quantizations_b200 import llama

    cfg = llama.LlamaConfig(hidden=1024, inter=2048, layers=2, heads=8, kv_heads=2, vocab=1000, max_len=64)
    dev_ = torch.device(DEV)
    m = llama.Llama(cfg, llama.linear4bit_factory(dev_, torch.bfloat16, "nf4"), dev_, torch.bfloat16)
    assert m.fuse_attn
    prompt = torch.arange(1, 12, device=dev_)
    m.forward(prompt, torch.arange(11, device=dev_))
    k0, v0 = m.k_cache.clone(), m.v_cache.clone()
    tok = torch.tensor([7], device=dev_)
    for step in range(4):
        pos = torch.tensor([11 + step], device=dev_)
        m.fuse_attn = False
        ref = m.forward(tok, pos).float()
        kr, vr = m.k_cache.clone(), m.v_cache.clone()
        m.k_cache.copy_(k0)
        m.v_cache.copy_(v0)
        m.fuse_attn = True
        got = m.forward(tok, pos).float()
        assert (got - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
        assert (m.k_cache.float() - kr.float()).abs().max().item() <= 2e-2 * kr.float().abs().max().item()
        assert (m.v_cache.float() - vr.float()).abs().max().item() <= 2e-2 * vr.float().abs().max().item()
        k0, v0 = m.k_cache.clone(), m.v_cache.clone()
        tok = got.argmax().view(1)


@pytest.mark.parametrize("M", [2, 4])
def test_small_batch_runs_one_decode_gemv_per_token(q, M):
    """2..4 tokens on a shape the one-pass small-batch kernel does not cover (rows not a multiple of 16) go through the decode GEMV
    row by row (modules.matmul_4bit): identical to calling the module on each token.  (Covered shapes: tests/test_tokens_gpu.py.)"""
    torch.manual_seed(11)
    lin = q.Linear4bit(1024, 776, bias=True, compute_dtype=torch.bfloat16, quant_type="nf4").to(DEV)
    lin.bias.data = torch.randn(776, device=DEV, dtype=torch.bfloat16)
    x = torch.randn(1, M, 1024, device=DEV, dtype=torch.bfloat16)
    n0 = q._lib.launch_count()
    y = lin(x)
    assert q._lib.launch_count() - n0 == M and y.shape == (1, M, 776)
    for m in range(M):
        assert torch.equal(y[:, m], lin(x[:, m:m + 1])[:, 0])


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_chained_gemvs_equal_separate_launches(q, dtype):
    """q4_gemv_4bit_chain: o_proj(+residual) -> norm + gate/up (grouped) -> SwiGLU + down(+residual) -> norm + next q/k/v in ONE
    persistent launch (grid barriers between the stages) gives bit-identical results to the four separate launches, on repeated
    calls (the barrier counter returns to zero) and under CUDA-graph replay."""
    from quantizations_b200 import graphs

    torch.manual_seed(21)
    dt = TDT[dtype]
    H, I = 1024, 2048
    mk = lambda n, k: q.Linear4bit(k, n, bias=False, compute_dtype=dt, quant_type="nf4").to(DEV)
    o, gate, up, down, qp, kp, vp = mk(H, H), mk(I, H), mk(I, H), mk(H, I), mk(H, H), mk(256, H), mk(256, H)
    gu, qkv = q.Linear4bitGroup([gate, up]), q.Linear4bitGroup([qp, kp, vp])
    ln2 = (1 + 0.1 * torch.randn(H, device=DEV)).to(dt)
    ln1 = (1 + 0.1 * torch.randn(H, device=DEV)).to(dt)
    a = torch.randn(1, 1, H, device=DEV, dtype=dt)
    h0 = torch.randn(1, 1, H, device=DEV, dtype=dt)

    def separate():
        h = h0.clone()
        q.gemv_4bit_fused(a, o.weight.data, o.weight.quant_state, residual=h, out=h)
        g_u = q.gemv_4bit_fused(h, None, group=gu, rms_weight=ln2)
        q.gemv_4bit_fused(g_u[..., I:], down.weight.data, down.weight.quant_state, gate=g_u[..., :I], residual=h, out=h)
        return h, q.gemv_4bit_fused(h, None, group=qkv, rms_weight=ln1)

    hs, qs = separate()
    h = h0.clone()
    g_u = torch.empty(1, 1, 2 * I, device=DEV, dtype=dt)
    out_qkv = torch.empty(1, 1, H + 512, device=DEV, dtype=dt)

    def chained():
        with q.gemv_4bit_chain() as ch:
            ch.add(a, o.weight.data, o.weight.quant_state, residual=h, out=h)
            ch.add(h, None, group=gu, rms_weight=ln2, out=g_u)
            ch.add(g_u[..., I:], down.weight.data, down.weight.quant_state, gate=g_u[..., :I], residual=h, out=h)
            ch.add(h, None, group=qkv, rms_weight=ln1, out=out_qkv)

    n0 = q._lib.launch_count()
    chained()
    assert q._lib.launch_count() - n0 == 1
    assert torch.equal(h, hs) and torch.equal(out_qkv, qs)
    for _ in range(3):
        h.copy_(h0)
        chained()
        assert torch.equal(h, hs) and torch.equal(out_qkv, qs)
    h.copy_(h0)
    g = graphs.capture(chained)   # warm-up calls + capture advance h: reset before every replay
    for _ in range(3):
        h.copy_(h0)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(h, hs) and torch.equal(out_qkv, qs)


@pytest.mark.parametrize("shape", [(4096, 4096), (1000, 512)])
@pytest.mark.parametrize("M", [2, 7, 16])
def test_small_batch_tcgen05_gemv_matches_per_token_gemv(q, shape, M):
    """q4_gemv_4bit_batch with Q4_GEMV_BATCH_TC5 (the tcgen05 decode kernel with the MMA's N columns as tokens): one pass over the packed weight for
    2..16 tokens gives what the per-token decode GEMV gives, bias included, ragged last row tile included."""
    N, K = shape
    torch.manual_seed(17)
    W = (torch.randn(N, K, device=DEV) * 0.02).to(torch.bfloat16)
    packed, st = q.quantize_4bit(W, quant_type="nf4")
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    x = torch.randn(1, M, K, device=DEV, dtype=torch.bfloat16)
    y = q.gemv_4bit_batch(x, packed, st, bias=bias, flags=q._lib.Q4_GEMV_BATCH_TC5)
    ref = torch.cat([q.gemv_4bit(x[:, m:m + 1], packed, state=st, bias=bias) for m in range(M)], dim=1)
    assert y.shape == (1, M, N)
    assert (y.float() - ref.float()).abs().max().item() <= 1e-2 * ref.float().abs().max().item()


def test_decode_kernels_do_not_write_outside_their_outputs(q):
    """Guard bands around every output of the decode kernels (mma.sync GEMV incl. ragged shapes and grouped launch, tcgen05 GEMV
    single / multi-token, chained launch, decode attention) must keep their sentinel after the calls."""
    dt = torch.bfloat16
    G = 256

    def guarded(n):
        buf = torch.full((n + 2 * G,), 7.75, device=DEV, dtype=dt)
        return buf, buf[G:G + n]

    def intact(buf, n):
        return bool((buf[:G] == 7.75).all()) and bool((buf[G + n:] == 7.75).all())

    torch.manual_seed(4)
    for N, K in ((512, 1024), (1000, 640), (264, 256)):
        lin = q.Linear4bit(K, N, bias=False, compute_dtype=dt, quant_type="nf4").to(DEV)
        st = lin.weight.quant_state
        x = torch.randn(1, 1, K, device=DEV, dtype=dt)
        buf, out = guarded(N)
        q.gemv_4bit(x, lin.weight.data, out=out.view(1, 1, N), state=st)
        assert intact(buf, N) and torch.isfinite(out.float()).all()
        if K % 256 == 0:
            M = 5
            buf, out = guarded(M * N)
            q.gemv_4bit_batch(torch.randn(1, M, K, device=DEV, dtype=dt), lin.weight.data, st, out=out.view(1, M, N))
            assert intact(buf, M * N) and torch.isfinite(out.float()).all()
    H, I = 512, 1024
    mk = lambda n, k: q.Linear4bit(k, n, bias=False, compute_dtype=dt, quant_type="nf4").to(DEV)
    o, gate, up, down = mk(H, H), mk(I, H), mk(I, H), mk(H, I)
    gu = q.Linear4bitGroup([gate, up])
    ln = torch.ones(H, device=DEV, dtype=dt)
    a = torch.randn(1, 1, H, device=DEV, dtype=dt)
    bh, h = guarded(H)
    bg, g_u = guarded(2 * I)
    h.copy_(torch.randn(H, device=DEV, dtype=dt))
    with q.gemv_4bit_chain() as ch:
        ch.add(a, o.weight.data, o.weight.quant_state, residual=h.view(1, 1, H), out=h.view(1, 1, H))
        ch.add(h.view(1, 1, H), None, group=gu, rms_weight=ln, out=g_u.view(1, 1, 2 * I))
        ch.add(g_u.view(1, 1, 2 * I)[..., I:], down.weight.data, down.weight.quant_state, gate=g_u.view(1, 1, 2 * I)[..., :I],
               residual=h.view(1, 1, H), out=h.view(1, 1, H))
    torch.cuda.synchronize()
    assert intact(bh, H) and intact(bg, 2 * I)
    nh, nkv, hd, L = 4, 2, 128, 16
    qkv_t = torch.randn(1, 1, (nh + 2 * nkv) * hd, device=DEV, dtype=dt)
    cos, sin = torch.randn(L, hd // 2, device=DEV, dtype=dt), torch.randn(L, hd // 2, device=DEV, dtype=dt)
    bk, kc = guarded(nkv * L * hd)
    bv, vc = guarded(nkv * L * hd)
    kc.zero_()
    vc.zero_()
    bo, ao = guarded(nh * hd)
    for p in (0, 7, L - 1):
        q.decode_attention(qkv_t, cos, sin, kc.view(nkv, L, hd), vc.view(nkv, L, hd), torch.tensor([p], device=DEV), nh, nkv,
                           out=ao.view(1, 1, nh * hd))
    torch.cuda.synchronize()
    assert intact(bk, nkv * L * hd) and intact(bv, nkv * L * hd) and intact(bo, nh * hd)


@pytest.mark.parametrize("quant_type,compress", [("nf4", True), ("fp4", False)])
def test_state_dict_round_trip_of_a_quantised_module(q, quant_type, compress):
    """ADVICE r1 / SURVEY f2: state_dict() of a quantised Linear4bit carries the statistics under bitsandbytes' key names, and
    load_state_dict() into a fresh module (Params4bit.from_prequantized underneath) gives the same forward outputs bit for bit;
    deepcopy and pickle keep the quantisation state as well."""
    import copy
    import io
    import pickle

    torch.manual_seed(3)
    dt = torch.bfloat16
    lin = q.Linear4bit(512, 256, bias=True, compute_dtype=dt, compress_statistics=compress, quant_type=quant_type).to(DEV)
    x1 = torch.randn(1, 1, 512, device=DEV, dtype=dt)
    x8 = torch.randn(1, 8, 512, device=DEV, dtype=dt)
    y1, y8 = lin(x1), lin(x8)
    sd = lin.state_dict()
    assert f"weight.quant_state.bitsandbytes__{quant_type}" in sd and "weight.absmax" in sd and "weight.quant_map" in sd
    assert ("weight.nested_absmax" in sd) == compress
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    sd2 = torch.load(buf, map_location="cpu")
    fresh = q.Linear4bit(512, 256, bias=True, compute_dtype=dt, compress_statistics=compress, quant_type=quant_type, device="meta")
    fresh.to_empty(device=DEV)
    missing, unexpected = fresh.load_state_dict(sd2, strict=True, assign=False)
    assert not missing and not unexpected
    assert fresh.weight.bnb_quantized and fresh.weight.quant_state.nested == compress
    assert torch.equal(fresh(x1), y1) and torch.equal(fresh(x8), y8)
    for clone in (copy.deepcopy(lin), pickle.loads(pickle.dumps(lin))):
        assert isinstance(clone.weight, q.Params4bit) and clone.weight.quant_state is not None
        assert torch.equal(clone(x1), y1) and torch.equal(clone(x8), y8)


def test_linear4bit_copies_and_pickles_after_a_decode_forward(q):
    """The cached launch descriptor of the single-vector forward (raw pointers in a ctypes struct) lives in the module's __dict__ but
    must not travel: a deep copy / pickle of a module that has already run drops it and rebuilds its own on first use; re-laying the
    weight (Linear4bitGroup) invalidates it."""
    import copy
    import pickle

    torch.manual_seed(2)
    lin = q.Linear4bit(512, 256, bias=True, compute_dtype=torch.bfloat16, quant_type="nf4").to(DEV)
    x = torch.randn(1, 1, 512, device=DEV, dtype=torch.bfloat16)
    y = lin(x)
    assert "_q4_decode" in lin.__dict__
    twin = copy.deepcopy(lin)
    assert twin.__dict__.get("_q4_decode") is None
    assert twin.weight.data_ptr() != lin.weight.data_ptr()
    assert torch.equal(twin(x), y)
    again = pickle.loads(pickle.dumps(lin))
    assert again.__dict__.get("_q4_decode") is None and torch.equal(again(x), y)
    other = q.Linear4bit(512, 128, bias=False, compute_dtype=torch.bfloat16, quant_type="nf4").to(DEV)
    lin.bias = None
    y0, y1 = lin(x), other(x)
    q.Linear4bitGroup([lin, other])
    assert "_q4_decode" not in lin.__dict__
    assert torch.equal(lin(x), y0) and torch.equal(other(x), y1)
