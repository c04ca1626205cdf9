"""Developer probe: a few launches of the fused dequantise + tcgen05 prefill GEMM (q4_gemm_4bit) for an ncu capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import quantizations_b200 as q  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, K = 14336, 4096
W = (torch.randn(N, K, device=dev) * 0.02).to(torch.float16)
packed, st = q.quantize_4bit(W, quant_type="fp4")
for M in (64, 512, 4096):
    x = torch.randn(1, M, K, device=dev, dtype=torch.float16)
    for _ in range(3):
        y = q.gemm_4bit(x, packed, st)
    torch.cuda.synchronize()
    ref = torch.nn.functional.linear(x, q.dequantize_4bit(packed, st).t().to(torch.float16))
    err = ((y.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
    print(f"M={M}: rel err vs dequantise + cuBLAS {err:.3e}")
    assert err < 2e-2
