"""oracle/ref_linear.py -- TEST / BENCH INFRASTRUCTURE ONLY.

A Linear module that runs the REFERENCE's kernels (oracle/_ref/kbkim_lib.so + ref_shim.so, compiled from /root/reference for
sm_100a) in the reference's own call sequence, so that the reference can be timed inside the same decode loop as the product.
It re-enacts reference modules.py:124-151 (Linear4bit.forward) -> core.py:426-504 (gemv_4bit): cast to the compute dtype,
dequantize_blockwise of the 8-bit absmax, torch `+= offset`, the GEMV kernel, cast back; prefill: dequantize_4bit
(core.py:581-634) + cast + F.linear (modules.py:63-64).  Everything launches on the legacy default stream like the reference.
The weights are quantised by the product's quantize_4bit (bit-identical packed format, tests/test_gpu_parity.py).
"""
import ctypes
import os
import sys

import torch
import torch.nn as nn

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_state = {}


def available() -> bool:
    return os.path.exists(os.path.join(_REF, "kbkim_lib.so")) and os.path.exists(os.path.join(_REF, "ref_shim.so"))


def _libs():
    if not _state:
        if _REF not in sys.path:
            sys.path.insert(0, _REF)
        import kbkim_lib

        shim = ctypes.CDLL(os.path.join(_REF, "ref_shim.so"))
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        for n in ("ref_gemv_bf16_async", "ref_gemv_fp16_async", "ref_gemv_fp32_async"):
            getattr(shim, n).argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32]
            getattr(shim, n).restype = None
        shim.ref_dequant_fp4_bf16.argtypes = [vp, vp, vp, i32, i32]
        _state.update(k=kbkim_lib, shim=shim)
    return _state["k"], _state["shim"]


class RefLinear4bit(nn.Module):
    """compute_dtype=None reproduces the as-shipped path (fp32 compute, the only GEMV instance the reference exports);
    torch.bfloat16 / torch.float16 use the instances it instantiates but does not export (ops.cu:176-177)."""

    def __init__(self, packed, quant_state, compute_dtype=None):
        super().__init__()
        self.packed, self.qs = packed, quant_state
        self.compute_dtype = compute_dtype if compute_dtype is not None else torch.float32
        nb = quant_state.absmax.numel()
        self.absmax_f32 = torch.empty(nb, device=packed.device, dtype=torch.float32)

    def forward(self, x):
        k, shim = _libs()
        qs = self.qs
        N, K = qs.shape
        inp_dtype = x.dtype
        x = x.to(self.compute_dtype)                                                     # modules.py:141-142
        s2 = qs.state2
        k.cdequantize_blockwise_fp32(s2.code.data_ptr(), qs.absmax.data_ptr(), s2.absmax.data_ptr(), self.absmax_f32.data_ptr(),
                                     s2.blocksize, qs.absmax.numel())                     # core.py:467 / :614
        self.absmax_f32 += qs.offset                                                      # core.py:468 / :615
        if x.numel() == x.shape[-1]:
            out = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=x.device)       # core.py:471-475
            fn = {torch.float32: shim.ref_gemv_fp32_async, torch.float16: shim.ref_gemv_fp16_async,
                  torch.bfloat16: shim.ref_gemv_bf16_async}[x.dtype]
            fn(N, 1, K, x.data_ptr(), self.packed.data_ptr(), self.absmax_f32.data_ptr(), qs.code.data_ptr(), out.data_ptr(),
               N, (K + 1) // 2, N, qs.blocksize)                                          # core.py:486-499
        else:
            w = torch.empty(N, K, dtype=torch.float16, device=x.device)                  # core.py:619 (quant_state.dtype = fp16)
            k.cdequantize_blockwise_fp16_fp4(0, self.packed.data_ptr(), self.absmax_f32.data_ptr(), w.data_ptr(), qs.blocksize,
                                             N * K)                                       # core.py:624 (FP4 tree only)
            out = torch.nn.functional.linear(x, w.to(x.dtype))                            # modules.py:64
        return out.to(inp_dtype)                                                          # modules.py:149


def ref_factory(device, dtype=torch.bfloat16, quant_type="fp4", compute_dtype=None, seed=0):
    """Same random weights as quantizations_b200.llama.linear4bit_factory, served by the reference's kernels."""
    import quantizations_b200 as q

    counter = [seed]

    def make(fin, fout, name):
        counter[0] += 1
        g = torch.Generator(device=device).manual_seed(counter[0])
        W = (torch.randn(fout, fin, device=device, dtype=torch.float32, generator=g) * 0.02).to(dtype)
        packed, qs = q.quantize_4bit(W, quant_type=quant_type)
        return RefLinear4bit(packed, qs, compute_dtype)

    return make
