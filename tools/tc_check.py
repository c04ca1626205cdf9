"""Developer tool: tcgen05 GEMV vs mma.sync GEMV on the same inputs, repeated calls; prints the rows that differ."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizations_b200 as q
from quantizations_b200 import _lib

dev = torch.device("cuda:0")
L = _lib.lib()
shapes = [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336), (6144, 4096), (28672, 4096), (1000, 512), (136, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
ws = torch.zeros(_lib.Q4_GEMV_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
for N, K in shapes:
    torch.manual_seed(0)
    W = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    packed, st = q.quantize_4bit(W, quant_type="nf4")
    stats = st.native_stats()
    lut = st.lut(torch.bfloat16)
    for rep in range(4):
        x = torch.randn(1, 1, K, device=dev, dtype=torch.bfloat16)
        o_tc = torch.full((N,), 7.0, device=dev, dtype=torch.bfloat16)
        o_mm = torch.full((N,), 9.0, device=dev, dtype=torch.bfloat16)
        for out, w in ((o_tc, ws), (o_mm, None)):
            f = _lib.GemvFused(x.data_ptr(), None, None, 0.0, packed.data_ptr(), ctypes.pointer(stats), None, None, 1, st.code.data_ptr(), None,
                               out.data_ptr(), N, K, 64, 2, 0, None, 0, lut.data_ptr(), None if w is None else w.data_ptr(), 0 if w is None else w.numel())
            rc = L.q4_gemv_4bit_fused(ctypes.byref(f), torch.cuda.current_stream().cuda_stream)
            assert rc == 0, rc
        torch.cuda.synchronize()
        d = (o_tc.float() - o_mm.float()).abs()
        tol = 2e-2 * o_mm.float().abs().max().item()
        bad = (d > tol).nonzero().view(-1)
        cnt = ws[: 65536].view(torch.int32)
        print(f"{N}x{K} rep {rep}: max diff {d.max().item():.4f} (tol {tol:.4f}) bad rows {bad.numel()}"
              + (f" first {bad[:6].tolist()} last {bad[-6:].tolist()}" if bad.numel() else "") + f"  counters nonzero {int((cnt != 0).sum())}")
