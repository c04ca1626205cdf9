"""Small run of every decode kernel for compute-sanitizer (memcheck): mma.sync GEMV (regular + ragged shapes, grouped, fused glue,
no table image), tcgen05 GEMV (1 and several tokens), chained launch, decode attention, quantize / dequantize, fused GEMM."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("Q4_GEMV_TC", "0")
import ctypes
import quantizations_b200 as q
from quantizations_b200 import _lib

dev = torch.device("cuda:0")
dt = torch.bfloat16
mk = lambda n, k: q.Linear4bit(k, n, bias=False, compute_dtype=dt, quant_type="nf4").to(dev)
for N, K in ((512, 1024), (1000, 640), (264, 256)):           # regular, ragged K (tail path), ragged rows
    lin = mk(N, K)
    x = torch.randn(1, 1, K, device=dev, dtype=dt)
    y = lin(x)
    st = lin.weight.quant_state
    stats = st.native_stats()
    # no table image: in-kernel build
    out = torch.empty(1, 1, N, device=dev, dtype=dt)
    rc = _lib.lib().q4_gemv_4bit(x.data_ptr(), lin.weight.data_ptr(), stats, st.code.data_ptr(), None, out.data_ptr(), N, K, 64, _lib.Q4_BF16, 0,
                                 None, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and torch.allclose(out.float(), y.float(), atol=1e-2, rtol=1e-2)
    if K % 256 == 0:
        yb = q.gemv_4bit_batch(torch.randn(1, 5, K, device=dev, dtype=dt), lin.weight.data, st)
        ws = torch.zeros(_lib.Q4_GEMV_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
        f = _lib.GemvFused(x.data_ptr(), None, None, 0.0, lin.weight.data_ptr(), ctypes.pointer(stats), None, None, 1, st.code.data_ptr(), None,
                           out.data_ptr(), N, K, 64, _lib.Q4_BF16, 0, None, 0, st.lut(dt).data_ptr(), ws.data_ptr(), ws.numel())
        assert _lib.lib().q4_gemv_4bit_fused(ctypes.byref(f), torch.cuda.current_stream().cuda_stream) == 0
    lin(torch.randn(1, 3, K, device=dev, dtype=dt))            # small batch: per-token GEMVs
    lin(torch.randn(1, 24, K, device=dev, dtype=dt))           # prefill: fused GEMM
H, I = 512, 1024
o, gate, up, down, qp, kp, vp = mk(H, H), mk(I, H), mk(I, H), mk(H, I), mk(H, H), mk(128, H), mk(128, H)
gu, qkv = q.Linear4bitGroup([gate, up]), q.Linear4bitGroup([qp, kp, vp])
ln = torch.ones(H, device=dev, dtype=dt)
a = torch.randn(1, 1, H, device=dev, dtype=dt)
h = torch.randn(1, 1, H, device=dev, dtype=dt)
g_u = torch.empty(1, 1, 2 * I, device=dev, dtype=dt)
oq = torch.empty(1, 1, H + 256, device=dev, dtype=dt)
with q.gemv_4bit_chain() as ch:
    ch.add(a, o.weight.data, o.weight.quant_state, residual=h, out=h)
    ch.add(h, None, group=gu, rms_weight=ln, out=g_u)
    ch.add(g_u[..., I:], down.weight.data, down.weight.quant_state, gate=g_u[..., :I], residual=h, out=h)
    ch.add(h, None, group=qkv, rms_weight=ln, out=oq)
nh, nkv, hd, L = 4, 2, 128, 32
qkv_t = torch.randn(1, 1, (nh + 2 * nkv) * hd, device=dev, dtype=dt)
cos, sin = torch.randn(L, hd // 2, device=dev, dtype=dt), torch.randn(L, hd // 2, device=dev, dtype=dt)
kc, vc = torch.zeros(nkv, L, hd, device=dev, dtype=dt), torch.zeros(nkv, L, hd, device=dev, dtype=dt)
for p in (0, 1, 9, L - 1):
    q.decode_attention(qkv_t, cos, sin, kc, vc, torch.tensor([p], device=dev), nh, nkv)
W = torch.randn(300, 192, device=dev, dtype=dt)
pk, st = q.quantize_4bit(W, quant_type="fp4")
q.dequantize_4bit(pk, st)
torch.cuda.synchronize()
print("sanitize_smoke ok")
