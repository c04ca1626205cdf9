"""oracle/q4_torch_cpu.py -- TEST INFRASTRUCTURE ONLY (CPU baseline leg of bench.py; checked against q4_oracle in tests).

The "torch-CPU dequantize -> matmul of the same packed format" that BASELINE.json's north_star asks to be timed on the GPU
box's host cores: unpack nibbles, look the 4-bit code up, multiply by the (double-quant decoded) per-block absmax, then
F.linear in fp32.  It restates reference core.py:611-631 (dequantize_4bit) + modules.py:64 (F.linear) with torch CPU ops;
it is a port, not the reference (the reference has no CPU path), and it is never used by the product.
"""
from __future__ import annotations

import torch


def decode_absmax(qabsmax: torch.Tensor, code2: torch.Tensor, absmax2: torch.Tensor, offset: torch.Tensor, blocksize2: int = 256):
    """reference core.py:467-468: dequantize_blockwise(absmax, state2) then += offset."""
    nb = qabsmax.numel()
    a2 = absmax2.repeat_interleave(blocksize2)[:nb]
    return code2[qabsmax.long()] * a2 + offset


def dequantize(packed: torch.Tensor, absmax: torch.Tensor, code: torch.Tensor, shape, blocksize: int = 64) -> torch.Tensor:
    """packed uint8 [(n+1)//2] -> float32 [N, K]; element 2i is the HIGH nibble of byte i (reference kernels.cu:558-559)."""
    n = shape[0] * shape[1]
    p = packed.reshape(-1)
    idx = torch.stack((p >> 4, p & 0xF), dim=1).reshape(-1)[:n].long()
    w = code[idx].reshape(-1, blocksize) * absmax[:, None]
    return w.reshape(shape)


def linear(x: torch.Tensor, packed, absmax, code, shape, blocksize: int = 64) -> torch.Tensor:
    """reference modules.py:64: F.linear(A, dequantize_4bit(B, qs).t()) in fp32."""
    return torch.nn.functional.linear(x.float(), dequantize(packed, absmax, code, shape, blocksize))
