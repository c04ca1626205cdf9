"""Developer tool: correctness + timing of the fused dequant + tcgen05 GEMM against a torch fp32 reference."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizations_b200 as q

dev = torch.device("cuda:0")
cases = [(16, 128, 64), (16, 256, 256), (33, 384, 512), (128, 4096, 4096), (300, 1024, 4096)] if len(sys.argv) < 2 else \
        [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
for dt in (torch.float16, torch.bfloat16):
    for (M, N, K) in cases:
        for qt, nested in (("fp4", True), ("nf4", False)):
            torch.manual_seed(M + N + K)
            W = (torch.randn(N, K, device=dev) * 0.02).to(dt)
            packed, st = q.quantize_4bit(W, quant_type=qt, compress_statistics=nested)
            Wd = q.dequantize_4bit(packed, st).t().float()
            X = torch.randn(M, K, device=dev, dtype=dt)
            bias = torch.randn(N, device=dev, dtype=dt)
            y = q.gemm_4bit(X, packed, st, bias=bias)
            torch.cuda.synchronize()
            ref = X.float() @ Wd.t() + bias.float()
            err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
            print(f"{str(dt)[6:]:9s} M={M:5d} N={N:5d} K={K:5d} {qt} nested={nested}: rel err {err:.2e}", "OK" if err < 1e-2 else "FAIL", flush=True)

if os.environ.get("GEMM_TIME"):
    def timeit(fn, iters=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters
    dt = torch.float16
    for (N, K) in ((4096, 4096), (14336, 4096), (4096, 14336)):
        W = (torch.randn(N, K, device=dev) * 0.02).to(dt)
        packed, st = q.quantize_4bit(W, quant_type="fp4")
        Wd = torch.empty(N, K, device=dev, dtype=dt)
        for M in (16, 64, 128, 256, 512, 1024, 2048, 4096):
            X = torch.randn(M, K, device=dev, dtype=dt)
            t_fused = timeit(lambda: q.gemm_4bit(X, packed, st))
            def ref_path():
                q.core._dequantize_4bit_into(packed, st, Wd)
                return torch.nn.functional.linear(X, Wd)
            t_ref = timeit(ref_path)
            t_mm = timeit(lambda: torch.nn.functional.linear(X, Wd))
            fl = 2.0 * M * N * K
            print(f"N={N:5d} K={K:5d} M={M:5d}: fused {t_fused:8.1f} us ({fl/t_fused/1e6:7.1f} TF/s)  dequant+cuBLAS {t_ref:8.1f} us  cuBLAS alone {t_mm:8.1f} us ({fl/t_mm/1e6:7.1f} TF/s)", flush=True)
