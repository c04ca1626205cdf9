"""Host cost per call of a single-vector Linear forward: nn.Linear (bf16, cuBLAS) against Linear4bit, on a shape whose kernel (4-5 us)
is shorter than the host path, so a loop of calls without synchronisation runs at the host's rate.  This is what HF's eager generate()
waits for at batch 1 (224 Linear calls per token)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import quantizations_b200 as q  # noqa: E402


def rate(mod, x, n=4000):
    with torch.no_grad():
        for _ in range(200):
            mod(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            mod(x)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


def main():
    dev = "cuda"
    x = torch.randn(1, 1, 4096, device=dev, dtype=torch.bfloat16)
    dense = torch.nn.Linear(4096, 4096, bias=False, device=dev, dtype=torch.bfloat16)
    lin = q.Linear4bit(4096, 4096, bias=False, compute_dtype=torch.bfloat16, quant_type="nf4").to(dev)
    for _ in range(3):
        print(f"nn.Linear {rate(dense, x):.2f} us/call   Linear4bit {rate(lin, x):.2f} us/call", flush=True)
    import cProfile
    import pstats

    pr = cProfile.Profile()
    with torch.no_grad():
        pr.enable()
        for _ in range(2000):
            lin(x)
        pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(12)


if __name__ == "__main__":
    main()
