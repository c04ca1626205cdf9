"""Developer probe for compute-sanitizer: small ring-kernel chains (plain, SwiGLU pair + residual through the tagged copy, a K % 512 == 256
shape) and the single-launch kernel on the same data; exits non-zero on a mismatch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import quantizations_b200 as q  # noqa: E402

dev = "cuda"
dt = torch.bfloat16
torch.manual_seed(0)
mk = lambda n, k: q.Linear4bit(k, n, bias=False, compute_dtype=dt, quant_type="nf4").to(dev)
H, I = 1024, 2048 + 256  # I = 2304: K % 512 == 256 for down_proj
o, gate, up, down = mk(H, H), mk(I, H), mk(I, H), mk(H, I)
gu = q.Linear4bitGroup([gate, up])
ln = (1 + 0.1 * torch.randn(H, device=dev)).to(dt)
a = torch.randn(1, 1, H, device=dev, dtype=dt)
h0 = torch.randn(1, 1, H, device=dev, dtype=dt)

h = h0.clone()
q.gemv_4bit_fused(a, o.weight.data, o.weight.quant_state, residual=h, out=h)
g_ref = q.gemv_4bit_fused(h, None, group=gu, rms_weight=ln)
q.gemv_4bit_fused(g_ref[..., I:], down.weight.data, down.weight.quant_state, gate=g_ref[..., :I], residual=h, out=h)
h_ref = h.clone()

h = h0.clone()
g_u = torch.empty(1, 1, 2 * I, device=dev, dtype=dt)
n0 = q._lib.launch_count()
for rep in range(2):
    h.copy_(h0)
    with q.gemv_4bit_chain() as ch:
        ch.add(a, o.weight.data, o.weight.quant_state, residual=h, out=h)
        ch.add(h, None, group=gu, rms_weight=ln, out=g_u)
        ch.add(g_u[..., I:], down.weight.data, down.weight.quant_state, gate=g_u[..., :I], residual=h, out=h)
    torch.cuda.synchronize()
    err = ((h.float() - h_ref.float()).abs().max() / h_ref.float().abs().max()).item()
    print("rep", rep, "launches", q._lib.launch_count() - n0, "rel err", err)
    assert err <= 1e-2
print("ok")
