"""Developer diagnostic: which k-blocks / row tiles of the fused GEMM are wrong (one-hot k-block masks on the activation)."""
import sys

import torch

sys.path.insert(0, ".")
import quantizations_b200 as q  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
M, N, K = 256, 1024, 4096
dt = torch.float16
W = (torch.randn(N, K, device=dev) * 0.02).to(dt)
packed, st = q.quantize_4bit(W, quant_type="fp4")
Wd = q.dequantize_4bit(packed, st).t().float()
X = torch.randn(M, K, device=dev, dtype=dt)
for rep in range(3):
    y = q.gemm_4bit(X, packed, st).float()
    ref = X.float() @ Wd.t()
    e = (y - ref).abs()
    print("full: rel err", (e.max() / ref.abs().max()).item(), "bad row tiles", [(i, round(e[:, i * 128:(i + 1) * 128].max().item(), 3)) for i in range(N // 128) if e[:, i * 128:(i + 1) * 128].max() > 0.05])
bad = {}
for j in range(K // 64):
    Xj = torch.zeros_like(X)
    Xj[:, j * 64:(j + 1) * 64] = X[:, j * 64:(j + 1) * 64]
    y = q.gemm_4bit(Xj, packed, st).float()
    ref = Xj.float() @ Wd.t()
    e = (y - ref).abs()
    if e.max() > 0.02:
        tiles = [i for i in range(N // 128) if e[:, i * 128:(i + 1) * 128].max() > 0.02]
        t0 = tiles[0]
        rows = slice(t0 * 128, (t0 + 1) * 128)
        xj = Xj[:, j * 64:(j + 1) * 64].float()
        counts = {}
        detail = []
        for r in range(t0 * 128, (t0 + 1) * 128):
            lab = "other"
            if (y[:, r] - ref[:, r]).abs().max() < 0.02:
                lab = "good"
            else:
                for d in (-9, -8, -6, -3, 3, 6, 8, 9, -1, 1, -2, 2, -4, 4, -5, 5, -7, 7):
                    jj = j + d
                    if 0 <= jj < K // 64 and (y[:, r] - xj @ Wd[r, jj * 64:(jj + 1) * 64]).abs().max() < 0.02:
                        lab = f"kb{d:+d}"
                        break
                if lab == "other" and y[:, r].abs().max() < 1e-3:
                    lab = "zero"
                detail.append((r - t0 * 128, lab))
            counts[lab] = counts.get(lab, 0) + 1
        guess = (counts, detail[:12])
        bad[j] = (tiles, guess)
print("bad k-blocks (row tiles, matching k-block offset):", bad)
