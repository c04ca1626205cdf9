"""Developer tool: small-batch (2..16 tokens) tcgen05 GEMV: correctness vs per-token GEMVs and timing vs the alternatives."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizations_b200 as q
dev = torch.device("cuda:0")
def timeit(fn, iters=20, reps=8):
    """device time per call: `reps` calls captured in one CUDA graph (no host launch cost), replayed `iters` times"""
    from quantizations_b200 import graphs
    g = graphs.capture(lambda: [fn() for _ in range(reps)])
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * reps)
for N, K in ((4096, 4096), (14336, 4096), (4096, 14336), (1000, 512)):
    W = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    packed, st = q.quantize_4bit(W, quant_type="nf4")
    bias = torch.randn(N, device=dev, dtype=torch.bfloat16)
    for M in (2, 4, 8, 16):
        x = torch.randn(1, M, K, device=dev, dtype=torch.bfloat16)
        y = q.gemv_4bit_batch(x, packed, st, bias=bias)
        ref = torch.cat([q.gemv_4bit(x[:, m:m + 1], packed, state=st, bias=bias) for m in range(M)], dim=1)
        err = ((y.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
        tb = timeit(lambda: q.gemv_4bit_batch(x, packed, st, bias=bias))
        tl = timeit(lambda: [q.gemv_4bit(x[:, m:m + 1], packed, state=st, bias=bias) for m in range(M)])
        tg = timeit(lambda: q.gemm_4bit(x, packed, st, bias=bias)) if K % 64 == 0 else float("nan")
        print(f"{N}x{K} M={M:2d}: err vs per-token GEMV {err:.1e}  batch {tb:7.1f} us  per-token loop {tl:7.1f} us  fused GEMM {tg:7.1f} us", flush=True)
