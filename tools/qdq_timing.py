"""Developer tool: HBM rate of the blockwise quantize / dequantize kernels on one GPU (algorithmic bytes).

    python tools/qdq_timing.py [--iters 40] [--pool 8]

Kernel-only timing through the C ABI (q4_quantize_blockwise_4bit / q4_dequantize_blockwise_4bit), CUDA events around
back-to-back launches that rotate through `pool` distinct matrices (inputs + outputs far larger than L2), so every byte
comes from / goes to HBM.  Also times the public quantize_4bit / dequantize_4bit calls (all launches of the recipe).
Not part of the product; bench.py is the contract benchmark.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantizations_b200 as q  # noqa: E402
from quantizations_b200 import _lib, core  # noqa: E402


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6531.6


def timed(fn, iters):
    for i in range(3):
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
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--shapes", default="4096x4096,14336x4096")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    L = _lib.lib()
    peak = peak_gbs()
    stream = lambda: torch.cuda.current_stream(dev).cuda_stream  # noqa: E731
    for shape in args.shapes.split(","):
        N, K = (int(v) for v in shape.split("x"))
        n = N * K
        for dtype in (torch.bfloat16, torch.float16):
            esz = 2
            torch.manual_seed(0)
            Ws = [(torch.randn(N, K, device=dev) * 0.02).to(dtype) for _ in range(args.pool)]
            for qt in ("nf4", "fp4"):
                packs, states = zip(*[q.quantize_4bit(W, quant_type=qt) for W in Ws])
                absmaxs = [torch.empty(n // 64, device=dev, dtype=torch.float32) for _ in range(args.pool)]
                outs = [torch.empty_like(p) for p in packs]
                deq = [torch.empty(N, K, device=dev, dtype=dtype) for _ in range(args.pool)]
                stats = [s.native_stats() for s in states]

                def kq(i):
                    j = i % args.pool
                    core.check(L.q4_quantize_blockwise_4bit(Ws[j].data_ptr(), absmaxs[j].data_ptr(), outs[j].data_ptr(), 64, n,
                                                            core._QUANT_CODE[qt], core._DTYPE_CODE[dtype], stream()), "quantize")

                def kd(i):
                    j = i % args.pool
                    core.check(L.q4_dequantize_blockwise_4bit(packs[j].data_ptr(), stats[j], deq[j].data_ptr(), 64, n,
                                                              core._QUANT_CODE[qt], core._DTYPE_CODE[dtype], stream()), "dequantize")

                def aq(i):
                    q.quantize_4bit(Ws[i % args.pool], quant_type=qt)

                def ad(i):
                    q.dequantize_4bit(packs[i % args.pool], states[i % args.pool])

                nb2 = -(-(n // 64) // 256)
                q_bytes = esz * n + n // 2 + 4 * (n // 64)                       # kernel: fp32 absmax out
                d_bytes = n // 2 + n // 64 + 4 * nb2 + 1092 + esz * n            # SURVEY 8d (nested statistics decoded in the kernel)
                tq, td = timed(kq, args.iters), timed(kd, args.iters)
                taq, tad = timed(aq, max(4, args.iters // 4)), timed(ad, max(4, args.iters // 4))
                assert all(torch.equal(o, p) for o, p in zip(outs, packs)), "kernel-only quantize differs from quantize_4bit"
                print(json.dumps({
                    "shape": shape, "dtype": str(dtype).split(".")[-1], "quant_type": qt,
                    "quantize_kernel_us": round(tq, 2), "quantize_GBs": round(q_bytes / tq / 1e3, 1),
                    "quantize_frac_of_measured_peak": round(q_bytes / tq / 1e3 / peak, 3),
                    "dequantize_kernel_us": round(td, 2), "dequantize_GBs": round(d_bytes / td / 1e3, 1),
                    "dequantize_frac_of_measured_peak": round(d_bytes / td / 1e3 / peak, 3),
                    "quantize_4bit_api_us": round(taq, 1), "dequantize_4bit_api_us": round(tad, 1),
                    "peak_GBs": peak, "pool": args.pool}), flush=True)
                del packs, states, absmaxs, outs, deq, stats
            del Ws
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
