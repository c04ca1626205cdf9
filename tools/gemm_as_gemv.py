import sys, os, torch
sys.path.insert(0, "/root/repo")
import quantizations_b200 as q
dev = torch.device("cuda:0")
for N, K in [(4096, 4096), (14336, 4096), (4096, 14336), (28672, 4096)]:
    W = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    mats = []
    packed, st = q.quantize_4bit(W, quant_type="nf4")
    nmat = max(2, min(48, (1 << 30) // (N * K // 2)))
    mats = [packed.clone() for _ in range(nmat)]
    for M in (1, 16):
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for i in range(5):
            q.gemm_4bit(x, mats[i % nmat], st, out=out)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for i in range(nmat):
                    q.gemm_4bit(x, mats[i], st, out=out)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e3 / (10 * nmat)
        B = N * K // 2 + N * K // 64
        print(f"{N}x{K} M={M}: {t:.2f} us  {B / t / 1e3:.0f} GB/s")
