"""Developer tool: phase timeline of the GEMV kernel (globaltimer stamps written by thread 0 of every CTA)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizations_b200 as q
from quantizations_b200 import _lib

N, K = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "14336x4096").split("x"))
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = _lib.lib()
L.q4_debug_set_gemv_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
W = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
packed, st = q.quantize_4bit(W, quant_type="nf4")
mats = [packed.clone() for _ in range(8)]
x = torch.randn(1, 1, K, device=dev, dtype=torch.bfloat16)
out = torch.empty(1, 1, N, device=dev, dtype=torch.bfloat16)
NL = 6
traces = [torch.zeros(1024 * 8 + 64, dtype=torch.int64, device=dev) for _ in range(NL)]  # kernels write blockIdx * 8 + slot, extras after
stream = torch.cuda.current_stream().cuda_stream
lut = None if os.environ.get("NO_LUT") else st.lut(torch.bfloat16)
stats = st.native_stats()
ws = torch.zeros(_lib.Q4_GEMV_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
fused = [_lib.GemvFused(x.data_ptr(), None, None, 0.0, m.data_ptr(), ctypes.pointer(stats), None, None, 1, st.code.data_ptr(), None,
                        out.data_ptr(), N, K, 64, 2, flags, None, 0, None if lut is None else lut.data_ptr(),
                        None if os.environ.get("NO_WS") else ws.data_ptr(), 0 if os.environ.get("NO_WS") else ws.numel()) for m in mats]
def launch(i, tr):
    L.q4_debug_set_gemv_trace(tr.data_ptr() if tr is not None else None)
    L.q4_gemv_4bit_fused(ctypes.byref(fused[i % 8]), stream)
for i in range(10):
    launch(i, None)
torch.cuda.synchronize()
# flush L2
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev); junk.fill_(1); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    stream = s.cuda_stream
    with torch.cuda.graph(g, stream=s):
        for i in range(NL):
            launch(i, traces[i])
g.replay(); torch.cuda.synchronize()
junk.fill_(2); torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
L.q4_debug_set_gemv_trace(None)
names = ["start", "issued", "waited", "x staged", "loop done", "end", "table ok", "tma sent"]
t0 = None
for i, tr in enumerate(traces):
    if i == NL - 1:
        ex = tr.cpu()
        print("   CTA 0 extra marks, us after its start:", [round((int(v) - int(ex[0])) / 1e3, 2) for v in ex[1024 * 8:1024 * 8 + 4]],
              " tma sent", round((int(ex[7]) - int(ex[0])) / 1e3, 2), " issued", round((int(ex[1]) - int(ex[0])) / 1e3, 2))
    t = tr.cpu()[:1024 * 8].view(1024, 8)
    t = t[t[:, 0] > 0]
    if t0 is None:
        t0 = int(t[:, 0].min())
    rel = (t - t0).float() / 1e3
    print(f"launch {i}: CTAs {t.shape[0]}")
    if i == NL - 1 and t.shape[0] > 148:
        lo, hi = rel[:148], rel[148:]
        print(f"   first 148 CTAs: start {lo[:, 0].median():.2f} x-staged {lo[:, 3].median():.2f} end {lo[:, 5].median():.2f} (max {lo[:, 5].max():.2f})"
              f" | rest: start {hi[:, 0].median():.2f} x-staged {hi[:, 3].median():.2f} end {hi[:, 5].median():.2f} (max {hi[:, 5].max():.2f})")
    for j, n in enumerate(names):
        col = rel[:, j]
        print(f"   {n:8s} min {col.min():8.2f}  median {col.median():8.2f}  max {col.max():8.2f} us")
