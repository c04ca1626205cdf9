"""The small-batch decode kernel (q4_gemv_4bit_batch -> csrc/q4_gemv_tokens.cu: 2..16 tokens in ONE pass over the packed weight,
the MMA's 8 B columns as tokens) against the fp64 truth of the oracle, against the per-token decode GEMV, and through
`Linear4bit.forward` / `matmul_4bit`, which dispatch it for 2..16 tokens.

Reference path replaced: modules.py:56-64 (anything but a single token -> dequantize_4bit + dense GEMM)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
TDT = {"bfloat16": torch.bfloat16, "float16": torch.float16}


@pytest.fixture(scope="module")
def q():
    import quantizations_b200 as q

    return q


def f32(t):
    return t.detach().float().cpu().numpy()


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
@pytest.mark.parametrize("quant_type,nested", [("nf4", True), ("nf4", False), ("fp4", True)])
@pytest.mark.parametrize("N,K,M", [(64, 256, 2), (48, 1280, 3), (1024, 4096, 8), (16, 14336, 5), (304, 2048, 16), (2064, 768, 9)])
def test_tokens_kernel_vs_fp64_truth(q, oracle, dtype, quant_type, nested, N, K, M):
    """tolerance: max|y - truth| <= 1e-2 * max|truth| per token (BASELINE north_star), and the tighter figure the batch-1 GEMV is
    held to (tests/test_gpu_parity.py::test_gemv_vs_fp64_truth): same arithmetic, the accumulator is just scaled per block."""
    rng = np.random.default_rng(N * 131 + K + M)
    dt = TDT[dtype]
    W = torch.from_numpy((rng.standard_normal((N, K)) * 0.02).astype(np.float32)).to(DEV).to(dt)
    packed, state = q.quantize_4bit(W, quant_type=quant_type, compress_statistics=nested)
    x = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(DEV).to(dt).reshape(1, M, K)
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(DEV).to(dt)
    n0 = q._lib.launch_count()
    y = q.gemv_4bit_batch(x, packed, state)
    assert q._lib.launch_count() - n0 == (M + 7) // 8  # one pass per 8 tokens
    yb = q.gemv_4bit_batch(x, packed, state, bias=bias)
    assert y.shape == (1, M, N) and y.dtype == dt
    st = oracle.quantize_4bit(f32(W), 64, quant_type, offset=float(state.offset.item()) if nested else None, compress_statistics=nested)
    tight = {"float16": 2e-3, "bfloat16": 8e-3}[dtype]
    for m in range(M):
        truth = oracle.gemv_4bit_f64(f32(x)[0, m], st["packed"], oracle.state_absmax(st), st["code"], N, K, st["blocksize"])
        scale = np.abs(truth).max()
        err = np.abs(f32(y)[0, m] - truth).max()
        assert err <= tight * scale, f"token {m}: max err {err:.3e} vs scale {scale:.3e}"
        errb = np.abs(f32(yb)[0, m] - (truth + f32(bias))).max()
        assert errb <= 1e-2 * np.abs(truth + f32(bias)).max()


@pytest.mark.parametrize("shape", [(4096, 4096), (14336, 4096), (4096, 14336), (6144, 4096)])
@pytest.mark.parametrize("M", [2, 4, 8, 13])
def test_tokens_kernel_matches_per_token_gemv_at_llama_sizes(q, shape, M):
    """Llama-3-8B shapes (q/k/v fused = 6144 rows): what one pass gives for M tokens is what M decode GEMVs give (different summation
    order: GEMV tolerance), and repeated calls are bit-identical (fixed-order combine, no atomics)."""
    N, K = shape
    torch.manual_seed(N + K + M)
    W = (torch.randn(N, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
    packed, st = q.quantize_4bit(W, quant_type="nf4")
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    y = q.gemv_4bit_batch(x, packed, st)
    ref = torch.cat([q.gemv_4bit(x[m:m + 1].view(1, 1, K), packed, state=st).view(1, N) for m in range(M)], dim=0)
    assert y.shape == (M, N)
    assert (y.float() - ref.float()).abs().max().item() <= 1e-2 * ref.float().abs().max().item()
    assert torch.equal(y, q.gemv_4bit_batch(x, packed, st))


@pytest.mark.parametrize("N,K,M", [(40000, 512, 3), (128256, 1024, 8), (20480, 2304, 16)])
def test_tokens_kernel_many_row_tiles_per_cta(q, N, K, M):
    """More row tiles per CTA than the kernel holds partial sums for at once (8): the pass loop, the ring restart between passes and the
    last, shorter pass (an lm_head-sized matrix: 128256 rows = 8016 tiles on 148 CTAs = 7 passes), K that leaves warps without chunks
    (512 = 2 chunks) and K whose chunks do not divide by the 16 warps (2304 = 9 chunks)."""
    torch.manual_seed(N + K + M)
    W = (torch.randn(N, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
    packed, st = q.quantize_4bit(W, quant_type="nf4")
    del W
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    y = q.gemv_4bit_batch(x, packed, st, bias=bias)
    ref = torch.cat([q.gemv_4bit(x[m:m + 1].view(1, 1, K), packed, state=st, bias=bias).view(1, N) for m in range(M)], dim=0)
    assert (y.float() - ref.float()).abs().max().item() <= 1e-2 * ref.float().abs().max().item()
    assert torch.equal(y, q.gemv_4bit_batch(x, packed, st, bias=bias))


def test_linear4bit_sends_2_to_16_tokens_through_one_pass(q):
    """modules.matmul_4bit: 2..16 tokens = one launch per 8 tokens; shapes the kernel does not cover (N % 16, K % 256) keep the old
    routes (per-token GEMVs up to 4 tokens, the fused GEMM beyond); results agree with the module applied token by token."""
    torch.manual_seed(11)
    lin = q.Linear4bit(1024, 768, bias=True, compute_dtype=torch.bfloat16, quant_type="nf4").to(DEV)
    lin.bias.data = torch.randn(768, device=DEV, dtype=torch.bfloat16)
    for M in (2, 4, 8, 9, 16):
        x = torch.randn(1, M, 1024, device=DEV, dtype=torch.bfloat16)
        n0 = q._lib.launch_count()
        y = lin(x)
        assert q._lib.launch_count() - n0 == (M + 7) // 8 and y.shape == (1, M, 768)
        ref = torch.cat([lin(x[:, m:m + 1]) for m in range(M)], dim=1)
        assert (y.float() - ref.float()).abs().max().item() <= 1e-2 * ref.float().abs().max().item()
    odd = q.Linear4bit(1024, 1000, bias=False, compute_dtype=torch.bfloat16, quant_type="nf4").to(DEV)  # 1000 rows: not whole 16-row tiles
    x = torch.randn(1, 3, 1024, device=DEV, dtype=torch.bfloat16)
    n0 = q._lib.launch_count()
    y = odd(x)
    assert q._lib.launch_count() - n0 == 3
    for m in range(3):
        assert torch.equal(y[:, m], odd(x[:, m:m + 1])[:, 0])


def test_tokens_kernel_under_graph_replay_and_guard_bands(q):
    """CUDA-graph capture (programmatic dependent launch on) and guard bands around the output."""
    from quantizations_b200 import graphs

    torch.manual_seed(3)
    N, K, M = 2048, 1024, 6
    lin = q.Linear4bit(K, N, bias=False, compute_dtype=torch.bfloat16, quant_type="nf4").to(DEV)
    st = lin.weight.quant_state
    G = 256
    buf = torch.full((M * N + 2 * G,), 7.75, device=DEV, dtype=torch.bfloat16)
    out = buf[G:G + M * N].view(1, M, N)
    x = torch.randn(1, M, K, device=DEV, dtype=torch.bfloat16)
    ref = q.gemv_4bit_batch(x, lin.weight.data, st)
    g = graphs.capture(lambda: q.gemv_4bit_batch(x, lin.weight.data, st, out=out, flags=q._lib.Q4_GEMV_PDL))
    for _ in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
    assert bool((buf[:G] == 7.75).all()) and bool((buf[G + M * N:] == 7.75).all())
