"""Timing of the small-batch decode paths on the Llama-3-8B shapes: the one-pass tokens kernel (q4_gemv_4bit_batch) against the batch-1
GEMV, M per-token GEMVs, the tcgen05 multi-token GEMV and the fused GEMM.  CUDA events around graph replays, rotating weights (> L2)."""
import sys

import torch

sys.path.insert(0, ".")
import quantizations_b200 as q  # noqa: E402

DEV = "cuda"
SHAPES = [(4096, 4096), (6144, 4096), (14336, 4096), (4096, 14336), (28672, 4096)]


def timeit(fn, n=20):
    """n launches captured in one CUDA graph (no host time between them), replayed 5 times"""
    from quantizations_b200 import graphs

    def body():
        for i in range(n):
            fn(i)

    g = graphs.capture(body)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n) * 1e3


def main():
    torch.manual_seed(0)
    peak = 6531.6
    for N, K in SHAPES:
        nmat = max(2, int(200e6 // (N * K // 2)) + 1)  # rotate > 126 MB of packed weights
        mats = []
        for _ in range(nmat):
            W = (torch.randn(N, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
            mats.append(q.quantize_4bit(W, quant_type="nf4"))
            del W
        bytes_alg = N * K // 2 + N * K // 64
        x1 = torch.randn(1, 1, K, device=DEV, dtype=torch.bfloat16)
        o1 = torch.empty(1, 1, N, device=DEV, dtype=torch.bfloat16)
        t1 = timeit(lambda i: q.gemv_4bit(x1, mats[i % nmat][0], o1, state=mats[i % nmat][1]))
        line = f"{N}x{K}: gemv M=1 {t1:.2f} us ({bytes_alg / t1 / 1e3 / peak:.3f});"
        for M in (2, 4, 8, 16):
            x = torch.randn(1, M, K, device=DEV, dtype=torch.bfloat16)
            o = torch.empty(1, M, N, device=DEV, dtype=torch.bfloat16)
            tb = timeit(lambda i: q.gemv_4bit_batch(x, mats[i % nmat][0], mats[i % nmat][1], out=o))
            line += f" tokens M={M} {tb:.2f} us ({bytes_alg / tb / 1e3 / peak:.3f}, x{tb / t1:.2f})"
            if M in (4, 16):
                ttc = timeit(lambda i: q.gemv_4bit_batch(x, mats[i % nmat][0], mats[i % nmat][1], out=o, flags=q._lib.Q4_GEMV_BATCH_TC5))
                tg = timeit(lambda i: q.gemm_4bit(x, mats[i % nmat][0], mats[i % nmat][1], out=o))
                line += f" [tc5 {ttc:.2f}, gemm {tg:.2f}]"
        print(line, flush=True)
        del mats


if __name__ == "__main__":
    main()
