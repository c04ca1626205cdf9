"""nn.Module layer -- host-side mirror of the reference's modules.py (Linear4bit, matmul_4bit).

reference: /root/reference/modules.py:28-64 (matmul_4bit), :67-151 (Linear4bit).  Inference only, like the reference.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib
from .core import Params4bit, QuantState, _dequantize_4bit_into, _on_device, fused_gemm_supported, gemm_4bit, gemv_4bit, gemv_4bit_batch


def matmul_4bit(A: torch.Tensor, B: torch.Tensor, quant_state: QuantState, out: torch.Tensor = None, bias=None,
                flags: int = _lib.Q4_GEMV_DEFAULT, prefetch: torch.Tensor = None):
    """A @ dequant(B)^T (+ bias).  reference modules.py:28-64.

    Decode (A is a single vector): one fused GEMV launch, bias included.
    Prefill: fp16/bf16 activations go to the fused dequantise + tcgen05 GEMM (`gemm_4bit`: the dense weight is never
    written to memory); other cases dequantise straight into A's dtype (one launch, no fp16 detour and no cast, unlike
    modules.py:64) and call F.linear.
    """
    assert quant_state is not None
    if A.numel() == A.shape[-1]:
        return gemv_4bit(A, B, out, state=quant_state, bias=bias, flags=flags, prefetch=prefetch)
    M = A.numel() // A.shape[-1]
    if M <= _SMALL_BATCH_GEMV_ROWS and A.dtype in (torch.float16, torch.bfloat16) and quant_state.blocksize == 64:
        n, k = quant_state.shape
        if k % 256 == 0 and n % 16 == 0 and (bias is None or bias.dtype == A.dtype):
            # 2..16 tokens (speculative / multi-sequence decode): ONE pass over the packed weight with the tokens on
            # the MMA's B columns (csrc/q4_gemv_tokens.cu) -- the price of a batch-1 GEMV for up to 8 tokens, of two for 9..16,
            # against 24-52 us for the fused GEMM at M=16 (profiles/r01b_prefill_gemm.txt)
            return gemv_4bit_batch(A, B, quant_state, bias=bias, out=out, flags=flags)
        if M <= 4:
            # shapes that kernel does not cover: one decode GEMV per token is still cheaper than any path that touches a dense weight
            A2 = A.reshape(M, A.shape[-1])
            if not A2.is_contiguous():
                A2 = A2.contiguous()
            res = out if out is not None else torch.empty(A.shape[:-1] + (quant_state.shape[0],), dtype=A.dtype, device=A.device)
            res2 = res.view(M, quant_state.shape[0])
            for m in range(M):
                gemv_4bit(A2[m:m + 1], B, res2[m:m + 1], state=quant_state, bias=bias, flags=flags)
            return res
    if _use_fused_gemm(A, quant_state, bias):
        return gemm_4bit(A, B, quant_state, bias=bias, out=out)
    W = torch.empty(quant_state.shape, dtype=A.dtype, device=A.device)
    _dequantize_4bit_into(B, quant_state, W)
    return torch.nn.functional.linear(A, W, bias)


_SMALL_BATCH_GEMV_ROWS = 16


def _use_fused_gemm(A, quant_state, bias) -> bool:
    """Prefill dispatch.  Q4_PREFILL=fused|cublas forces a path; `auto` (default) takes the fused tcgen05 kernel where it
    measured faster than dequantise + cuBLAS on B200 (profiles/r02_prefill_gemm.txt): up to 64 tokens on every Llama-3 shape,
    up to 256 tokens when out_features >= in_features (4096x4096, 14336x4096, q/k/v), up to 512 on the 14336-wide ones; beyond
    that the dense GEMM dominates and cuBLAS' 2-CTA kernels are still ahead of this kernel's single-CTA pipeline."""
    if not fused_gemm_supported(A, quant_state) or not (bias is None or bias.dtype == A.dtype):
        return False
    mode = os.environ.get("Q4_PREFILL", "auto")
    if mode == "fused":
        return True
    if mode == "cublas":
        return False
    M = A.numel() // A.shape[-1]
    n, k = quant_state.shape
    if M <= 64:
        return True
    if n < k:
        return False
    return M <= 256 or (M <= 512 and n * k >= 32 * 1024 * 1024)


class Linear4bit(nn.Linear):
    """Linear layer over a 4-bit blockwise-quantised weight; drop-in for the reference's Linear4bit
    (modules.py:67-151) and, through it, for bitsandbytes.nn.Linear4bit as HF transformers constructs it:
    Linear4bit(in, out, bias, compute_dtype, compress_statistics=..., quant_type=..., quant_storage=...).

    quant_type: "fp4" (reference) or "nf4".  compress_statistics=False keeps fp32 absmax (the reference accepts the
    flag and ignores it, modules.py:80).
    """

    def __init__(
        self,
        input_features,
        output_features,
        bias=False,
        compute_dtype=None,
        compress_statistics=True,
        quant_type="fp4",
        quant_storage=torch.uint8,
        device=None,
    ):
        super().__init__(input_features, output_features, bias, device)
        self.weight = Params4bit(
            self.weight.data,
            requires_grad=False,
            quant_type=quant_type,
            quant_storage=quant_storage,
            module=self,
            compress_statistics=compress_statistics,
        )
        self.compute_dtype = compute_dtype
        self.compute_type_is_set = False
        self.quant_state = None
        self.quant_storage = quant_storage
        # decode-launch hints (new; see include/quantizations_b200.h): programmatic dependent launch is always safe --
        # the kernel reads x and writes its output only after the preceding kernel has completed
        self.gemv_flags = _lib.Q4_GEMV_PDL
        self.prefetch_next = None  # (packed weight, in_features) of the Linear that runs next: its first tiles are pulled into L2

    # ---- state dict: the packed weight travels with its statistics under bitsandbytes' key names (`weight.absmax`,
    # `weight.quant_map`, `weight.nested_*`, `weight.quant_state.bitsandbytes__nf4` ...), so that state_dict() / load_state_dict()
    # / save_pretrained round-trip a quantised module.  The reference has no such hooks: its state dict carries the packed bytes only.

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        qs = getattr(self.weight, "quant_state", None)
        if qs is not None:
            for k, v in qs.as_dict(packed=True).items():
                destination[prefix + "weight." + k] = v if keep_vars else v.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        wkey = prefix + "weight"
        stat_keys = [k for k in state_dict if k.startswith(wkey + ".")]
        if stat_keys and wkey in state_dict:
            packed = state_dict[wkey]
            stats = {k[len(wkey) + 1:]: state_dict[k] for k in stat_keys}
            dev = packed.device if packed.is_cuda else (self.weight.device if self.weight.is_cuda else torch.device("cuda"))
            self.weight = Params4bit.from_prequantized(packed, stats, device=dev, module=self)
            self.quant_state = self.weight.quant_state
            _drop_decode_cache(self)
            state_dict = {k: v for k, v in state_dict.items() if k not in stat_keys}
            state_dict[wkey] = self.weight.data  # what the generic loader copies into the (already rebuilt) parameter
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def set_compute_type(self, x):
        """reference modules.py:112-122: fp32 / bf16 inputs set the compute dtype; fp16 keeps the configured one."""
        if x.dtype in [torch.float32, torch.bfloat16]:
            self.compute_dtype = x.dtype

    def forward(self, x: torch.Tensor):
        """reference modules.py:124-151"""
        if not self.compute_type_is_set:
            self.set_compute_type(x)
            self.compute_type_is_set = True

        inp_dtype = x.dtype
        if self.compute_dtype is not None and x.dtype != self.compute_dtype:
            x = x.to(self.compute_dtype)
        bias = self.bias
        if bias is not None and bias.dtype != x.dtype:
            bias = bias.to(x.dtype)
        weight = self.weight
        if weight.quant_state is None:
            raise RuntimeError("Linear4bit weight is not quantized yet: move the module to a CUDA device first")
        if x.numel() == x.shape[-1] and x.dtype in _FAST_DTYPES and x.is_contiguous() and x.shape[-1] == self.in_features:
            out = self._decode(x, weight, bias)
            if out is not None:
                return out if out.dtype == inp_dtype else out.to(inp_dtype)
        out = matmul_4bit(x, weight.data, bias=bias, quant_state=weight.quant_state, flags=self.gemv_flags,
                          prefetch=self.prefetch_next)
        return out if out.dtype == inp_dtype else out.to(inp_dtype)


_FAST_DTYPES = (torch.float16, torch.bfloat16)


def _none():
    return None


class _DecodeLaunch:
    """Cached launch descriptor (q4_gemv_fused_t) of a Linear4bit's single-vector forward.  Lives in the module's __dict__ (one
    dictionary probe per call) but never travels with it: it deep-copies and pickles as None (ctypes descriptors hold raw pointers)."""
    __slots__ = ("weight", "qs", "wptr", "dtype", "flags", "f", "ref", "keep", "fn", "N", "pf")

    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_none, ())


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream  # the current stream's handle as an int, without building a Stream object
    _cur_device = torch._C._cuda_getDevice
except AttributeError:  # a torch build without the private accessors
    _raw_stream = lambda idx: torch.cuda.current_stream(idx).cuda_stream  # noqa: E731
    _cur_device = torch.cuda.current_device


def _linear4bit_decode(self, x, weight, bias):
    """Single-vector forward with a cached launch descriptor (q4_gemv_fused_t): everything that does not change between calls --
    weight, statistics, tables -- is bound once, so a call costs one allocation, three pointer stores and the launch.  Under HF
    generate() the host side of 224 Linear calls per token is what batch-1 decode waits for (tests/perf/hf_generate_tps.py), so
    the per-call Python is kept to identity checks.  Same launch as core.gemv_4bit."""
    qs = weight.quant_state
    c = self.__dict__.get("_q4_decode")
    if (c is None or c.weight is not weight or c.qs is not qs or c.dtype != x.dtype or c.flags != self.gemv_flags
            or c.wptr != weight.data_ptr()):
        if qs.blocksize != 64:
            return None
        import ctypes

        from .core import _DTYPE_CODE, _ws_args

        stats, lut = qs.native_stats(), qs.lut(x.dtype)
        c = _DecodeLaunch()
        c.weight, c.qs, c.wptr, c.dtype, c.flags = weight, qs, weight.data_ptr(), x.dtype, self.gemv_flags
        c.f = _lib.GemvFused(None, None, None, 0.0, weight.data_ptr(), ctypes.pointer(stats), None, None, 1, qs.code.data_ptr(), None, None,
                             qs.shape[0], qs.shape[1], 64, _DTYPE_CODE[x.dtype], self.gemv_flags, None, 0, lut.data_ptr(),
                             *_ws_args(x.device))
        c.ref, c.keep, c.fn, c.N, c.pf = ctypes.byref(c.f), (stats, lut), _lib.lib().q4_gemv_4bit_fused, qs.shape[0], None
        self.__dict__["_q4_decode"] = c
    f = c.f
    out = torch.empty(x.shape[:-1] + (c.N,), dtype=x.dtype, device=x.device)
    f.x, f.out, f.bias = x.data_ptr(), out.data_ptr(), (None if bias is None else bias.data_ptr())
    nxt = self.prefetch_next
    if nxt is not c.pf:
        if nxt is not None:
            t, k = nxt if isinstance(nxt, tuple) else (nxt, 0)
            f.prefetch, f.prefetch_bytes, f.prefetch_K = t.data_ptr(), t.numel() * t.element_size(), k
        else:
            f.prefetch, f.prefetch_bytes, f.prefetch_K = None, 0, 0
        c.pf = nxt
    idx = x.device.index
    if idx == _cur_device():
        rc = c.fn(c.ref, _raw_stream(idx))
    else:
        with torch.cuda.device(idx):
            rc = c.fn(c.ref, _raw_stream(idx))
    if rc:
        _lib.check(rc, "gemv_4bit")
    return out


def _drop_decode_cache(module):
    """After anything that re-lays a module's weight or statistics in place (grouping, load_state_dict)."""
    module.__dict__.pop("_q4_decode", None)


class Linear4bitGroup(nn.Module):
    """Several quantised Linear4bit layers that consume the same activation (q/k/v, gate/up), served by ONE decode launch.

    The members' packed weights and statistics are re-laid back to back in one allocation (each member keeps working on its
    own: its `weight.data` / QuantState tensors become views of the shared buffers), and a single-vector forward runs
    q4_gemv_4bit_grouped over the concatenated rows: one prologue (table build, activation staging) instead of one per
    member, and a grid sized for the sum of the rows.  Anything else (prefill, unsupported dtype) falls back to the members.
    Requirements: all quantised, same in_features / blocksize 64 / quant_type / nesting, no bias, and for nested statistics
    N_i * K / 64 a multiple of 256 so that no second-level block straddles two members.
    """

    def __init__(self, linears, swiglu: bool = False):
        """swiglu=True (two equally shaped members: gate, up): a SECOND copy of the pair is laid out interleaved in chunks of four
        rows (rows 8t..8t+3 = gate rows 4t..4t+3, rows 8t+4..8t+7 = up rows 4t..4t+3; packed bytes, 8-bit absmax codes and
        second-level absmax alike), which lets `forward_swiglu` return silu(gate(x)) * up(x) from ONE launch: every CTA then owns
        both halves of its outputs (q4_gemv_fused_t flags & Q4_GEMV_SWIGLU).  The members keep their own storage for the other
        paths (prefill), so the pair costs twice its packed size."""
        super().__init__()
        import ctypes

        from ._lib import AbsmaxStats

        self.swiglu = bool(swiglu)
        if self.swiglu:
            self._init_swiglu(linears)
            return
        self.members = nn.ModuleList(linears)
        first = linears[0].weight.quant_state
        K = linears[0].in_features
        nested = first.nested
        for lin in linears:
            qs = lin.weight.quant_state
            if qs is None or lin.in_features != K or qs.blocksize != 64 or qs.quant_type != first.quant_type or qs.nested != nested:
                raise ValueError("Linear4bitGroup members must be quantised alike and share in_features")
            if lin.bias is not None:
                raise ValueError("Linear4bitGroup does not support biases")
            if nested and (lin.out_features * K // 64) % qs.state2.blocksize:
                raise ValueError("nested statistics would straddle two members")
        if len(linears) > 4 or K % 64:
            raise ValueError("at most 4 members, in_features a multiple of 64")
        dev = linears[0].weight.device
        self.in_features, self.splits = K, [lin.out_features for lin in linears]
        self.out_features = sum(self.splits)
        self.packed = torch.cat([lin.weight.data.reshape(-1) for lin in linears]).reshape(-1, 1)
        self.absmax = torch.cat([lin.weight.quant_state.absmax for lin in linears])
        self.absmax2 = torch.cat([lin.weight.quant_state.state2.absmax for lin in linears]) if nested else None
        o_p = o_a = o_a2 = 0
        for lin in linears:  # members become views of the shared buffers
            qs = lin.weight.quant_state
            n_p, n_a = lin.weight.data.numel(), qs.absmax.numel()
            lin.weight.data = self.packed[o_p:o_p + n_p]
            qs.absmax = self.absmax[o_a:o_a + n_a]
            if nested:
                n_a2 = qs.state2.absmax.numel()
                qs.state2.absmax = self.absmax2[o_a2:o_a2 + n_a2]
                qs.state2._stats = None
                o_a2 += n_a2
            qs._stats = None
            _drop_decode_cache(lin)
            o_p, o_a = o_p + n_p, o_a + n_a
        self.code, self.blocksize, self.nested = first.code, 64, nested
        if nested:
            self._stats = AbsmaxStats(None, self.absmax.data_ptr(), first.state2.code.data_ptr(), self.absmax2.data_ptr(),
                                      first.offset.data_ptr(), int(first.state2.blocksize))
            self._offsets = (ctypes.c_void_p * len(linears))(*[lin.weight.quant_state.offset.data_ptr() for lin in linears])
        else:
            self._stats = AbsmaxStats(self.absmax.data_ptr(), None, None, None, None, 0)
            self._offsets = None
        ends, acc = [], 0
        for n in self.splits:
            acc += n
            ends.append(acc)
        self._row_end = (ctypes.c_int * len(linears))(*ends)
        self.gemv_flags = _lib.Q4_GEMV_PDL
        self.prefetch_next = None
        self.device_ = dev

    def _init_swiglu(self, linears):
        import ctypes

        from ._lib import AbsmaxStats

        if len(linears) != 2:
            raise ValueError("a SwiGLU group has exactly two members (gate, up)")
        gate, up = linears
        self.members = nn.ModuleList(linears)
        qg, qu = gate.weight.quant_state, up.weight.quant_state
        K, N = gate.in_features, gate.out_features
        if qg is None or qu is None or up.in_features != K or up.out_features != N or qg.blocksize != 64 or qu.blocksize != 64 \
                or qg.quant_type != qu.quant_type or qg.nested != qu.nested or gate.bias is not None or up.bias is not None:
            raise ValueError("SwiGLU group members must be quantised alike, equally shaped, without bias")
        nested = qg.nested
        per4 = 4 * K // 64                       # 64-wide blocks in a 4-row chunk
        if N % 4 or K % 64 or (nested and (per4 % qg.state2.blocksize or qg.state2.blocksize != qu.state2.blocksize)):
            raise ValueError("interleaving needs out_features % 4 == 0 and 4 rows to hold whole second-level blocks (in_features % 4096 == 0)")

        def weave(a, b, width):                  # [N/4 chunks, width] each -> chunks alternating a, b
            return torch.stack((a.reshape(N // 4, width), b.reshape(N // 4, width)), dim=1).reshape(-1)

        self.in_features, self.splits, self.out_features = K, [N, N], 2 * N
        self.packed = weave(gate.weight.data.reshape(-1), up.weight.data.reshape(-1), 4 * K // 2).reshape(-1, 1)
        self.absmax = weave(qg.absmax, qu.absmax, per4)
        self.code, self.blocksize, self.nested = qg.code, 64, nested
        if nested:
            self.absmax2 = weave(qg.state2.absmax, qu.state2.absmax, per4 // qg.state2.blocksize)
            self._stats = AbsmaxStats(None, self.absmax.data_ptr(), qg.state2.code.data_ptr(), self.absmax2.data_ptr(),
                                      qg.offset.data_ptr(), int(qg.state2.blocksize))
            self._offsets = (ctypes.c_void_p * 2)(qg.offset.data_ptr(), qu.offset.data_ptr())
        else:
            self.absmax2 = None
            self._stats = AbsmaxStats(self.absmax.data_ptr(), None, None, None, None, 0)
            self._offsets = None
        self._row_end = (ctypes.c_int * 2)(N, 2 * N)
        self.gemv_flags = _lib.Q4_GEMV_PDL
        self.prefetch_next = None
        self.device_ = gate.weight.device

    def forward_swiglu(self, x: torch.Tensor, out: torch.Tensor = None, prefetch_override=None, **kw) -> torch.Tensor:
        """[.., K] single vector -> silu(gate(x)) * up(x)  [.., N], one launch (optionally with the RMSNorm fused in: rms_weight=)."""
        from .core import gemv_4bit_fused

        if not self.swiglu:
            raise ValueError("not a SwiGLU group")
        return gemv_4bit_fused(x, None, group=self, out=out, flags=self.gemv_flags | _lib.Q4_GEMV_SWIGLU,
                               prefetch=prefetch_override if prefetch_override is not None else self.prefetch_next, **kw)

    def lut(self, dtype):
        """Decode table image shared by the members (same code tables by construction)."""
        return self.members[0].weight.quant_state.lut(dtype)

    def forward_fused(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """[.., K] single vector -> [.., sum N_i] (concatenated member outputs)."""
        from .core import gemv_4bit_fused

        if self.swiglu:
            raise ValueError("a SwiGLU group's rows are interleaved: use forward_swiglu")

        return gemv_4bit_fused(x, None, group=self, out=out, flags=self.gemv_flags, prefetch=self.prefetch_next)

    def forward(self, x: torch.Tensor):
        """Returns one output per member, like calling them in turn."""
        if self.swiglu:  # the interleaved copy only serves forward_swiglu
            return tuple(lin(x) for lin in self.members)
        if x.numel() == x.shape[-1] and x.dtype in (torch.float16, torch.bfloat16) and x.is_contiguous():
            return self.forward_fused(x).split(self.splits, dim=-1)
        return tuple(lin(x) for lin in self.members)


Linear4bit._decode = _linear4bit_decode
