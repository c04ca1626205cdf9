"""The persistent ring GEMV (q4_gemv_4bit_ring, csrc/q4_gemv_ring.cuh) against separate launches of the single-GEMV kernel on the
same inputs: the decoder-layer chain at Llama-3-8B sizes (o + residual -> norm + gate/up -> SwiGLU + down + residual -> norm + q/k/v),
plain chains, fp16 / non-nested statistics, CUDA-graph replay, and the refusals that make callers fall back.

Reference path replaced: core.py:467-499 (three launches per Linear) x the layer's Linears, modules.py:124-151."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
TDT = {"bfloat16": torch.bfloat16, "float16": torch.float16}


@pytest.fixture(scope="module")
def q():
    import quantizations_b200 as q

    return q


def _close(a, b, tol=1e-2):
    a, b = a.float(), b.float()
    return (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1e-6)


def _ring_launch(q, stages):
    """stages: list of GemvFused structs (already chained through their pointers); returns the ABI's return code"""
    L = q._lib.lib()
    ws = q.core.ring_workspace(DEV)
    arr = (q._lib.GemvFused * len(stages))(*stages)
    return L.q4_gemv_4bit_ring(arr, len(stages), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
@pytest.mark.parametrize("compress", [True, False])
def test_ring_decoder_layer_chain_matches_separate_launches(q, dtype, compress):
    """Llama-3-8B sizes.  The ring kernel splits rows across CTAs along K and sums the per-(row, k tile) partials in k order, the
    arithmetic of the single-launch kernel: results must agree to the GEMV tolerance (they are in fact expected bit-identical)."""
    torch.manual_seed(5)
    dt = TDT[dtype]
    H, I, KV = 4096, 14336, 1024
    mk = lambda n, k: q.Linear4bit(k, n, bias=False, compute_dtype=dt, compress_statistics=compress, quant_type="nf4").to(DEV)
    o, gate, up, down, qp, kp, vp = mk(H, H), mk(I, H), mk(I, H), mk(H, I), mk(H, H), mk(KV, H), mk(KV, H)
    gu, qkv = q.Linear4bitGroup([gate, up]), q.Linear4bitGroup([qp, kp, vp])
    ln2 = (1 + 0.1 * torch.randn(H, device=DEV)).to(dt)
    ln1 = (1 + 0.1 * torch.randn(H, device=DEV)).to(dt)
    a = torch.randn(1, 1, H, device=DEV, dtype=dt)
    h0 = torch.randn(1, 1, H, device=DEV, dtype=dt)

    h = h0.clone()
    q.gemv_4bit_fused(a, o.weight.data, o.weight.quant_state, residual=h, out=h)
    g_ref = q.gemv_4bit_fused(h, None, group=gu, rms_weight=ln2)
    q.gemv_4bit_fused(g_ref[..., I:], down.weight.data, down.weight.quant_state, gate=g_ref[..., :I], residual=h, out=h)
    h_ref, qkv_ref = h, q.gemv_4bit_fused(h, None, group=qkv, rms_weight=ln1)

    h = h0.clone()
    g_u = torch.empty(1, 1, 2 * I, device=DEV, dtype=dt)
    out_qkv = torch.empty(1, 1, H + 2 * KV, device=DEV, dtype=dt)
    stages = []
    q.gemv_4bit_fused(a, o.weight.data, o.weight.quant_state, residual=h, out=h, _defer=stages)
    q.gemv_4bit_fused(h, None, group=gu, rms_weight=ln2, out=g_u, _defer=stages)
    q.gemv_4bit_fused(g_u[..., I:], down.weight.data, down.weight.quant_state, gate=g_u[..., :I], residual=h, out=h, _defer=stages)
    q.gemv_4bit_fused(h, None, group=qkv, rms_weight=ln1, out=out_qkv, _defer=stages)
    n0 = q._lib.launch_count()
    for rep in range(3):  # the epochs of the exchange advance from call to call
        h.copy_(h0)
        g_u.fill_(float("nan"))
        out_qkv.fill_(float("nan"))
        rc = _ring_launch(q, [f for f, _ in stages])
        assert rc == 0, q._lib.lib().q4_error_string(rc)
        torch.cuda.synchronize()
        assert _close(g_u, g_ref) and _close(h, h_ref) and _close(out_qkv, qkv_ref)
    assert q._lib.launch_count() - n0 == 3
    exact = torch.equal(g_u, g_ref) and torch.equal(h, h_ref) and torch.equal(out_qkv, qkv_ref)
    print("ring vs single-launch kernel bit-identical:", exact)


def test_ring_plain_chain_and_eight_stages(q):
    """Eight plainly chained stages (x of stage i+1 = the first K outputs of stage i), weights scaled so that magnitudes stay O(1)."""
    torch.manual_seed(9)
    dt = torch.bfloat16
    shapes = [(6144, 4096), (4096, 4096), (8192, 4096), (4096, 8192), (4096, 4096), (2048, 4096), (1024, 2048), (512, 1024)]
    lins = []
    for n, k in shapes:
        lin = q.Linear4bit(k, n, bias=False, compute_dtype=dt, quant_type="nf4", device="meta")
        W = (torch.randn(n, k, device=DEV) / k ** 0.5).to(dt)
        lin.weight = q.Params4bit(W, requires_grad=False, quant_type="nf4", module=lin).to(DEV)
        lins.append(lin)
    x0 = torch.randn(1, 1, 4096, device=DEV, dtype=dt)
    ref, x = [], x0
    for lin, (n, k) in zip(lins, shapes):
        y = q.gemv_4bit(x[..., :k].contiguous(), lin.weight.data, state=lin.weight.quant_state)
        ref.append(y)
        x = y
    outs = [torch.full((1, 1, n), float("nan"), device=DEV, dtype=dt) for n, _ in shapes]
    stages, x = [], x0
    for lin, (n, k), out in zip(lins, shapes, outs):
        q.gemv_4bit_fused(x[..., :k] if x is not x0 else x, lin.weight.data, lin.weight.quant_state, out=out, _defer=stages)
        x = out
    rc = _ring_launch(q, [f for f, _ in stages])
    assert rc == 0, q._lib.lib().q4_error_string(rc)
    torch.cuda.synchronize()
    for i, (y, r) in enumerate(zip(outs, ref)):
        assert _close(y, r), f"stage {i}"


def test_ring_refuses_what_it_does_not_support(q):
    """Ragged shapes, a later stage reading an earlier out in an unsupported way, too many stages: Q4_ERR_SHAPE and nothing
    launched -- the Python chain then falls back to the older launches with the same results."""
    dt = torch.bfloat16
    torch.manual_seed(2)
    mk = lambda n, k: q.Linear4bit(k, n, bias=False, compute_dtype=dt, quant_type="nf4").to(DEV)
    x = torch.randn(1, 1, 1024, device=DEV, dtype=dt)
    a, c, d = mk(1000, 1024), mk(1024, 1024), mk(512, 1024)  # a: rows % 32 != 0
    q.core.ring_workspace(DEV)
    n0 = q._lib.launch_count()
    st = []
    q.gemv_4bit_fused(x, a.weight.data, a.weight.quant_state, _defer=st)
    assert _ring_launch(q, [f for f, _ in st]) == q._lib.Q4_ERR_SHAPE
    st = []
    y1 = q.gemv_4bit_fused(x, c.weight.data, c.weight.quant_state, _defer=st)
    q.gemv_4bit_fused(x, d.weight.data, d.weight.quant_state, _defer=st)
    q.gemv_4bit_fused(y1, d.weight.data, d.weight.quant_state, _defer=st)  # reads stage 0's out, but is not its successor
    assert _ring_launch(q, [f for f, _ in st]) == q._lib.Q4_ERR_SHAPE
    assert q._lib.launch_count() == n0
    # the context manager hides the difference
    with q.gemv_4bit_chain() as ch:
        y = ch.add(x, a.weight.data, a.weight.quant_state)
    assert _close(y, q.gemv_4bit(x, a.weight.data, state=a.weight.quant_state), 0.0) or torch.equal(y, q.gemv_4bit(x, a.weight.data, state=a.weight.quant_state))


def test_older_chained_launch_still_matches_through_the_abi(q):
    """q4_gemv_4bit_chain (grid-barrier kernel) is no longer what gemv_4bit_chain() issues -- the ring kernel took over, single
    launches are the fallback -- but the export stays: called directly it must still equal the separate launches bit for bit."""
    torch.manual_seed(4)
    dt = torch.bfloat16
    mk = lambda n, k: q.Linear4bit(k, n, bias=False, compute_dtype=dt, quant_type="nf4").to(DEV)
    a, b = mk(1024, 1024), mk(512, 1024)
    x = torch.randn(1, 1, 1024, device=DEV, dtype=dt)
    y1 = q.gemv_4bit(x, a.weight.data, state=a.weight.quant_state)
    y2 = q.gemv_4bit(y1, b.weight.data, state=b.weight.quant_state)
    st = []
    o1 = q.gemv_4bit_fused(x, a.weight.data, a.weight.quant_state, _defer=st)
    o2 = q.gemv_4bit_fused(o1, b.weight.data, b.weight.quant_state, _defer=st)
    bar = torch.zeros(64, dtype=torch.int32, device=DEV)
    arr = (q._lib.GemvFused * 2)(*[f for f, _ in st])
    rc = q._lib.lib().q4_gemv_4bit_chain(arr, 2, bar.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(o1, y1) and torch.equal(o2, y2) and int(bar.abs().sum()) == 0


def test_ring_ragged_k_tile_pair_and_tagged_residual(q):
    """K % 512 == 256 (tensor-parallel shard sizes): the last k tile is a single TMA box and the stale half of the slot meets
    zero activations; together with the SwiGLU pair and a residual read through the tagged copy of an earlier stage."""
    dt = torch.bfloat16
    torch.manual_seed(0)
    mk = lambda n, k: q.Linear4bit(k, n, bias=False, compute_dtype=dt, quant_type="nf4").to(DEV)
    H, I = 1024, 2048 + 256
    o, gate, up, down = mk(H, H), mk(I, H), mk(I, H), mk(H, I)
    gu = q.Linear4bitGroup([gate, up])
    ln = (1 + 0.1 * torch.randn(H, device=DEV)).to(dt)
    a = torch.randn(1, 1, H, device=DEV, dtype=dt)
    h0 = torch.randn(1, 1, H, device=DEV, dtype=dt)
    h = h0.clone()
    q.gemv_4bit_fused(a, o.weight.data, o.weight.quant_state, residual=h, out=h)
    g_ref = q.gemv_4bit_fused(h, None, group=gu, rms_weight=ln)
    q.gemv_4bit_fused(g_ref[..., I:], down.weight.data, down.weight.quant_state, gate=g_ref[..., :I], residual=h, out=h)
    h_ref = h.clone()
    h = h0.clone()
    g_u = torch.full((1, 1, 2 * I), float("nan"), device=DEV, dtype=dt)
    stages = []
    q.gemv_4bit_fused(a, o.weight.data, o.weight.quant_state, residual=h, out=h, _defer=stages)
    q.gemv_4bit_fused(h, None, group=gu, rms_weight=ln, out=g_u, _defer=stages)
    q.gemv_4bit_fused(g_u[..., I:], down.weight.data, down.weight.quant_state, gate=g_u[..., :I], residual=h, out=h, _defer=stages)
    for rep in range(2):
        h.copy_(h0)
        assert _ring_launch(q, [f for f, _ in stages]) == 0
        torch.cuda.synchronize()
        assert _close(h, h_ref) and _close(g_u, g_ref)


def test_ring_thirty_two_stages_with_residuals_and_too_distant_residual(q):
    """32 stages in ONE launch (the exchange buffers are reused modulo 8): a residual stream updated in place every second stage
    (`bias` = the out of the stage two back, read from its tagged copy), checked against 32 single launches; twice, so that a
    buffer is reused across launches as well.  A residual more than six stages back would find its exchange buffer overwritten:
    refused with Q4_ERR_SHAPE."""
    torch.manual_seed(13)
    dt = torch.bfloat16
    H = 2048
    lins = []
    for i in range(32):
        lin = q.Linear4bit(H, H, bias=False, compute_dtype=dt, quant_type="nf4", device="meta")
        W = (torch.randn(H, H, device=DEV) * (0.5 / H ** 0.5)).to(dt)
        lin.weight = q.Params4bit(W, requires_grad=False, quant_type="nf4", module=lin).to(DEV)
        lins.append(lin)
    x0 = torch.randn(1, 1, H, device=DEV, dtype=dt)

    def reference():
        h = x0.clone()      # residual stream
        t = x0.clone()      # what the next stage consumes
        for i, lin in enumerate(lins):
            if i % 2 == 0:  # plain stage
                t = q.gemv_4bit(t, lin.weight.data, state=lin.weight.quant_state)
            else:           # h += W t, and the next stage consumes h
                q.gemv_4bit_fused(t, lin.weight.data, lin.weight.quant_state, residual=h, out=h)
                t = h
        return h.clone()

    ref = reference()
    h = torch.empty_like(x0)
    tmp = [torch.empty(1, 1, H, device=DEV, dtype=dt) for _ in range(16)]

    def build():
        stages, t = [], x0
        for i, lin in enumerate(lins):
            if i % 2 == 0:
                q.gemv_4bit_fused(t, lin.weight.data, lin.weight.quant_state, out=tmp[i // 2], _defer=stages)
                t = tmp[i // 2]
            else:
                q.gemv_4bit_fused(t, lin.weight.data, lin.weight.quant_state, residual=h, out=h, _defer=stages)
                t = h
        return stages

    stages = build()
    for _ in range(2):
        h.copy_(x0)
        n0 = q._lib.launch_count()
        rc = _ring_launch(q, [f for f, _ in stages])
        assert rc == 0, q._lib.lib().q4_error_string(rc)
        assert q._lib.launch_count() - n0 == 1
        torch.cuda.synchronize()
        assert _close(h, ref, 2e-2)  # 32 dependent stages: rounding flips propagate
    # residual eight stages back: stage 9's bias = stage 1's out
    far = []
    t = x0
    outs = [torch.empty(1, 1, H, device=DEV, dtype=dt) for _ in range(10)]
    for i in range(10):
        kw = {"residual": outs[1]} if i == 9 else {}
        q.gemv_4bit_fused(t, lins[i].weight.data, lins[i].weight.quant_state, out=outs[i], _defer=far, **kw)
        t = outs[i]
    n0 = q._lib.launch_count()
    assert _ring_launch(q, [f for f, _ in far]) == q._lib.Q4_ERR_SHAPE
    assert q._lib.launch_count() == n0
