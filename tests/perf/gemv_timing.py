"""Developer tool: per-shape GEMV timing on one GPU (product vs the reference kernel when oracle/_ref is present).

    python tests/perf/gemv_timing.py [--dtype bf16] [--quant nf4] [--pool-mb 1024] [--iters 400]

Weights rotate through a pool larger than L2 so every byte comes from HBM; times are CUDA-event averages over
back-to-back launches on the current stream.  Not part of the product; bench.py is the contract benchmark.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import quantizations_b200 as q  # noqa: E402
from quantizations_b200 import _lib  # noqa: E402

SHAPES = [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336)]


def algo_bytes(N, K, xbytes=2, nested=True):
    n = N * K
    b = n // 2 + xbytes * K + xbytes * N + 64
    b += (n // 64 + 4 * -(-n // 16384) + 1024 + 4) if nested else 4 * n // 64
    return b


NSTREAMS = 1


def time_graph(fn, nlaunch, replays=20):
    """Capture `nlaunch` back-to-back launches into one CUDA graph and time its replays: no host launch cost.
    With NSTREAMS > 1 the launches are dealt round-robin to parallel capture branches (independent launches)."""
    s = torch.cuda.Stream()
    side = [torch.cuda.Stream() for _ in range(NSTREAMS - 1)]
    with torch.cuda.stream(s):
        for i in range(nlaunch):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for t in side:
                t.wait_stream(s)
            for i in range(nlaunch):
                k = i % NSTREAMS
                if k == 0:
                    fn(i)
                else:
                    with torch.cuda.stream(side[k - 1]):
                        fn(i)
            for t in side:
                s.wait_stream(t)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (replays * nlaunch)


def time_fn(fn, iters, warmup=20):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--quant", default="nf4")
    ap.add_argument("--pool-mb", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=400)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--only", default="", help="NxK: time just this shape")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-ws", action="store_true", help="no split-K workspace: the mma.sync kernel instead of the tcgen05 one")
    ap.add_argument("--no-lut", action="store_true", help="no prebuilt table image: the kernel builds its table per launch")
    ap.add_argument("--next", action="store_true", help="hint the next matrix of the pool for L2 prefetch")
    ap.add_argument("--streams", type=int, default=1, help="parallel capture branches (independent launches)")
    a = ap.parse_args()
    global NSTREAMS
    NSTREAMS = a.streams
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    shim = None
    if not a.no_ref and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_shim.so")):
        shim = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "ref_shim.so"))
    print(f"device {torch.cuda.get_device_name(0)}  dtype {a.dtype} quant {a.quant}  peak {peak} GB/s  flags {a.flags}")
    shapes = [tuple(int(v) for v in a.only.split("x"))] if a.only else SHAPES
    for N, K in shapes:
        per = N * K // 2
        nmat = max(2, min(64, a.pool_mb * (1 << 20) // per))
        torch.manual_seed(0)
        W = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16 if dt == torch.float32 else dt)
        packed0, st0 = q.quantize_4bit(W, quant_type=a.quant)
        mats = []
        for i in range(nmat):
            p = packed0.clone() if i else packed0
            if i:
                p.view(-1)[i::97] ^= 0x5A  # distinct contents, same statistics
            mats.append(p)
        x = torch.randn(1, 1, K, device=dev, dtype=dt)
        out = torch.empty(1, 1, N, device=dev, dtype=dt)
        outs2 = [torch.empty(1, 1, N, device=dev, dtype=dt) for _ in range(4)]
        L = _lib.lib()
        stats = st0.native_stats()
        stream = torch.cuda.current_stream().cuda_stream
        dcode = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[dt]
        ptrs = [m.data_ptr() for m in mats]

        lut = None if (a.no_lut or dt == torch.float32) else st0.lut(dt)
        ws = torch.zeros(_lib.Q4_GEMV_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
        fused = [_lib.GemvFused(x.data_ptr(), None, None, 0.0, ptrs[i], ctypes.pointer(stats), None, None, 1, st0.code.data_ptr(), None,
                                outs2[i % NSTREAMS].data_ptr(), N, K, 64, dcode, a.flags, ptrs[(i + 1) % nmat] if a.next else None,
                                per if a.next else 0, None if lut is None else lut.data_ptr(),
                                None if a.no_ws else ws.data_ptr(), 0 if a.no_ws else ws.numel()) for i in range(nmat)]

        def ours(i):
            if dt == torch.float32:
                L.q4_gemv_4bit(x.data_ptr(), ptrs[i % nmat], stats, st0.code.data_ptr(), None, outs2[i % NSTREAMS].data_ptr(), N, K, 64,
                               dcode, a.flags, None, 0, torch.cuda.current_stream().cuda_stream)
            else:
                L.q4_gemv_4bit_fused(ctypes.byref(fused[i % nmat]), torch.cuda.current_stream().cuda_stream)

        # accuracy against an fp32 matmul of the dequantised weight
        ours(0)
        truth = (x.double().view(1, K) @ q.dequantize_4bit(mats[0], st0).double()).view(-1)
        err = ((outs2[0].double().view(-1) - truth).abs().max() / truth.abs().max()).item()

        def ours_py(i):
            q.gemv_4bit(x, mats[i % nmat], out=out, state=st0)

        tp = time_fn(ours_py, a.iters)
        t = time_fn(ours, a.iters) if a.no_graph else time_graph(ours, nmat * 2)
        B = algo_bytes(N, K, x.element_size())
        line = f"{N:6d}x{K:<6d} ours(graph) {t:8.2f} us  {B / t / 1e3:8.1f} GB/s ({B / t / 1e3 / peak * 100:5.1f}% of measured peak)  err {err:.1e}  eager core.gemv_4bit {tp:8.2f} us"
        if shim is not None:
            absmax = (q.dequantize_blockwise(st0.absmax, st0.state2) + st0.offset).contiguous()
            fn = {torch.float32: shim.ref_gemv_fp32, torch.float16: shim.ref_gemv_fp16, torch.bfloat16: shim.ref_gemv_bf16}[dt]
            # the shim synchronises after every call (it is a test shim), so time the reference with one event pair per call
            args = lambda i: (N, 1, K, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(ptrs[i % nmat]),
                              ctypes.c_void_p(absmax.data_ptr()), ctypes.c_void_p(st0.code.data_ptr()),
                              ctypes.c_void_p(out.data_ptr()), N, (K + 1) // 2, N, 64)
            for i in range(5):
                fn(*args(i))
            ts = []
            for i in range(40):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(*args(i))
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            tr = ts[len(ts) // 2]
            line += f"  | reference kernel (median single launch) {tr:8.2f} us {B / tr / 1e3:8.1f} GB/s"
        print(line, flush=True)


if __name__ == "__main__":
    main()
