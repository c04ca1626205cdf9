"""Developer probe: the fused prefill GEMM at decode-sized batches (M = 4, 16) for an ncu capture and a timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import quantizations_b200 as q  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, K = 14336, 4096
Ws = [(torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16) for _ in range(6)]
qs = [q.quantize_4bit(W, quant_type="nf4") for W in Ws]
for M in (4, 16):
    x = torch.randn(1, M, K, device=dev, dtype=torch.bfloat16)
    for i in range(6):
        y = q.gemm_4bit(x, *qs[i % 6])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60):
        y = q.gemm_4bit(x, *qs[i % 6])
    e1.record()
    torch.cuda.synchronize()
    print(f"M={M}: {e0.elapsed_time(e1) / 60 * 1e3:.1f} us per GEMM")
