"""Minimal batch-1 Llama decoder built on Linear4bit -- the end-to-end harness for BASELINE.json config 3/5.

Not a model zoo: it exists so that the 4-bit Linear hot path can be measured inside a real decode loop (KV cache, RoPE, GQA
attention, RMSNorm, SwiGLU) without HF generate()'s host overhead, and so that one decode step can be captured into a CUDA
graph (every kernel of this library runs on the current stream; the reference's run on the legacy default stream and cannot be
captured).  Weights are random-init (no checkpoints offline); shapes follow Llama-3.

`linear_factory(in_features, out_features, name) -> nn.Module` decides what each projection is: quantizations_b200.Linear4bit
(default), a dense nn.Linear (the "HF native" baseline), or the reference-kernel module of oracle/ (tests/bench only).
Tensor parallelism: q/k/v/gate/up column-parallel, o/down row-parallel with an all-reduce after each.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .core import argmax, decode_attention, gemv_4bit_chain, gemv_4bit_fused


@dataclass
class LlamaConfig:
    hidden: int = 4096
    inter: int = 14336
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    vocab: int = 128256
    rope_theta: float = 500000.0
    eps: float = 1e-5
    max_len: int = 256

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


LLAMA3_8B = LlamaConfig()
LLAMA3_70B = LlamaConfig(hidden=8192, inter=28672, layers=80, heads=64, kv_heads=8)


class DecoderLayer(nn.Module):
    def __init__(self, cfg: LlamaConfig, layer: int, make: Callable, tp: int):
        super().__init__()
        h, hd = cfg.hidden, cfg.head_dim
        self.nh, self.nkv, self.hd = cfg.heads // tp, cfg.kv_heads // tp, hd
        self.q_proj = make(h, self.nh * hd, f"L{layer}.q_proj")
        self.k_proj = make(h, self.nkv * hd, f"L{layer}.k_proj")
        self.v_proj = make(h, self.nkv * hd, f"L{layer}.v_proj")
        self.o_proj = make(self.nh * hd, h, f"L{layer}.o_proj")
        self.gate_proj = make(h, cfg.inter // tp, f"L{layer}.gate_proj")
        self.up_proj = make(h, cfg.inter // tp, f"L{layer}.up_proj")
        self.down_proj = make(cfg.inter // tp, h, f"L{layer}.down_proj")
        self.ln1 = nn.Parameter(torch.ones(h))
        self.ln2 = nn.Parameter(torch.ones(h))
        # siblings that read the same activation are served by one grouped decode launch when they are Linear4bit
        self.qkv = self.gate_up = self.gate_up_sw = None
        try:
            from .modules import Linear4bit, Linear4bitGroup

            if all(isinstance(m, Linear4bit) for m in (self.q_proj, self.k_proj, self.v_proj, self.gate_proj, self.up_proj)):
                self.qkv = Linear4bitGroup([self.q_proj, self.k_proj, self.v_proj])
                self.gate_up = Linear4bitGroup([self.gate_proj, self.up_proj])
                if os.environ.get("Q4_SWIGLU", "0") == "1":
                    # opt-in: silu(gate) * up straight out of the gate/up launch, from an interleaved SECOND copy of the pair.
                    # Measured 571 vs 566 tok/s on Llama-3-8B: not worth twice the gate/up memory by default.
                    try:
                        self.gate_up_sw = Linear4bitGroup([self.gate_proj, self.up_proj], swiglu=True)
                    except ValueError:
                        self.gate_up_sw = None
        except ValueError:
            self.qkv = self.gate_up = self.gate_up_sw = None


class Llama(nn.Module):
    def __init__(self, cfg: LlamaConfig, linear_factory: Callable, device, dtype=torch.bfloat16, tp: int = 1, group=None):
        super().__init__()
        self.cfg, self.tp, self.group, self.dtype = cfg, tp, group, dtype
        self.fuse_glue = True  # fold RMSNorm / SwiGLU / residual adds into the decode GEMV launches (Linear4bit layers only)
        self.fused_ar = None   # tp.FusedAllReduce: the row-parallel all-reduce inside the GEMV epilogue instead of NCCL
        self.fuse_swiglu = True  # decode: silu(gate) * up in the gate/up launch's epilogue (layers whose gate_up_sw group exists)
        self.fuse_attn = cfg.head_dim == 128  # decode: RoPE + KV append + attention as one launch (q4_decode_attention)
        # the decode step's attention launch reads `pos`, cos / sin and the cache rows of EARLIER steps while the q/k/v GEMV before
        # it is still running (only the new token's q/k/v depend on that GEMV)
        from . import _lib as _l
        self.attn_flags = _l.Q4_GEMV_PDL | _l.Q4_ATTN_EARLY_CACHE
        self.chain = False     # decode: o -> gate/up -> down -> next layer's q/k/v as ONE persistent launch of the ring kernel
                               # (q4_gemv_4bit_ring through core.gemv_4bit_chain; 603 vs 566 tok/s on Llama-3-8B);
                               # bench.py switches it on where the ring kernel covers the model's shapes
        g = torch.Generator(device=device).manual_seed(1234)
        self.embed = (torch.randn(cfg.vocab, cfg.hidden, device=device, dtype=torch.float32, generator=g) * 0.02).to(dtype)
        self.lm_head = (torch.randn(cfg.vocab, cfg.hidden, device=device, dtype=torch.float32, generator=g) * 0.02).to(dtype)
        self.norm = torch.ones(cfg.hidden, device=device, dtype=dtype)
        self.layers = nn.ModuleList([DecoderLayer(cfg, i, linear_factory, tp) for i in range(cfg.layers)])
        for layer in self.layers:
            layer.ln1.data = layer.ln1.data.to(device=device, dtype=dtype)
            layer.ln2.data = layer.ln2.data.to(device=device, dtype=dtype)
        hd = cfg.head_dim
        inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, device=device, dtype=torch.float32) / hd))
        ang = torch.arange(cfg.max_len, device=device, dtype=torch.float32)[:, None] * inv[None, :]
        self.cos, self.sin = ang.cos().to(dtype), ang.sin().to(dtype)      # [max_len, hd/2]
        nkv = cfg.kv_heads // tp
        self.k_cache = torch.zeros(cfg.layers, nkv, cfg.max_len, hd, device=device, dtype=dtype)
        self.v_cache = torch.zeros(cfg.layers, nkv, cfg.max_len, hd, device=device, dtype=dtype)
        self.positions = torch.arange(cfg.max_len, device=device)

    def _rope(self, x, cos, sin):  # x [T, heads, hd]; cos/sin [T, hd/2]  (HF "rotate_half" convention)
        x1, x2 = x[..., : self.cfg.head_dim // 2], x[..., self.cfg.head_dim // 2:]
        c, s = cos[:, None, :], sin[:, None, :]
        return torch.cat((x1 * c - x2 * s, x2 * c + x1 * s), dim=-1)

    @staticmethod
    def _hint(nxt):
        """Exact L2 prefetch hint for a decode launch: (packed weight, in_features) of the Linear4bit / group that runs next."""
        if nxt is None:
            return None
        return (nxt.packed if hasattr(nxt, "packed") else nxt.weight.data, nxt.in_features)

    def _allreduce(self, t):
        if self.tp > 1:
            torch.distributed.all_reduce(t, group=self.group)
        return t

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        """tokens [T] int64, pos [T] int64 (absolute positions, consecutive).  Returns logits of the LAST token [vocab].
        T == 1 is the decode step (static shapes: capturable in a CUDA graph with device-resident `tokens` / `pos`)."""
        cfg, T = self.cfg, tokens.shape[0]
        x = self.embed.index_select(0, tokens).unsqueeze(0)                 # [1, T, h]
        all_fused = (T == 1 and self.fuse_glue and self.fuse_attn and all(L.qkv is not None for L in self.layers)
                     and (self.tp == 1 or self.fused_ar is not None))
        if not all_fused:  # the fused decode step reads cos / sin / cache positions inside q4_decode_attention
            cos, sin = self.cos.index_select(0, pos), self.sin.index_select(0, pos)
            # causal mask over the whole static cache: key j visible to query i iff j <= pos[i]
            mask = self.positions[None, :] <= pos[:, None]                  # [T, max_len]
        if all_fused and self.chain:
            return self._decode_chained(x, pos)
        for li, L in enumerate(self.layers):
            fused = L.qkv is not None and T == 1 and self.fuse_glue
            h = None if fused else F.rms_norm(x, (cfg.hidden,), L.ln1, cfg.eps)
            if fused and self.fuse_attn:
                qkv = gemv_4bit_fused(x, None, group=L.qkv, rms_weight=L.ln1, rms_eps=cfg.eps, prefetch=self._hint(L.o_proj))
                a = decode_attention(qkv, self.cos, self.sin, self.k_cache[li], self.v_cache[li], pos, L.nh, L.nkv, flags=self.attn_flags)
            elif fused:  # RMSNorm folded into the grouped q/k/v launch: the norm kernel and its round trip disappear
                q, k, v = gemv_4bit_fused(x, None, group=L.qkv, rms_weight=L.ln1, rms_eps=cfg.eps).split(L.qkv.splits, dim=-1)
            elif L.qkv is not None and T == 1:
                q, k, v = L.qkv(h)
            else:
                q, k, v = L.q_proj(h), L.k_proj(h), L.v_proj(h)
            if not (fused and self.fuse_attn):
                q, k, v = q.reshape(T, L.nh, L.hd), k.reshape(T, L.nkv, L.hd), v.reshape(T, L.nkv, L.hd)
                q, k = self._rope(q, cos, sin), self._rope(k, cos, sin)
                self.k_cache[li].index_copy_(1, pos, k.transpose(0, 1))
                self.v_cache[li].index_copy_(1, pos, v.transpose(0, 1))
                a = F.scaled_dot_product_attention(q.transpose(0, 1).unsqueeze(0), self.k_cache[li].unsqueeze(0),
                                                   self.v_cache[li].unsqueeze(0), attn_mask=mask, enable_gqa=True)  # [1, nh, T, hd]
                a = a.squeeze(0).transpose(0, 1).reshape(1, T, L.nh * L.hd)
            if fused and (self.tp == 1 or self.fused_ar is not None):
                # o_proj adds the residual stream in its epilogue; norm folded into gate/up; SwiGLU folded into down_proj's
                # activation staging, residual again in its epilogue: four launches for the layer's seven Linears + glue
                qs = L.o_proj.weight.quant_state
                nxt = self.layers[li + 1].qkv if li + 1 < len(self.layers) else None
                sw = L.gate_up_sw if self.fuse_swiglu else None
                x = gemv_4bit_fused(a, L.o_proj.weight.data, qs, residual=x, allreduce=self.fused_ar,
                                    prefetch=self._hint(sw if sw is not None else L.gate_up))
                if sw is not None:  # SwiGLU in the gate/up epilogue: down_proj stages a plain vector
                    hmid = sw.forward_swiglu(x, rms_weight=L.ln2, rms_eps=cfg.eps, prefetch_override=self._hint(L.down_proj))
                    x = gemv_4bit_fused(hmid, L.down_proj.weight.data, L.down_proj.weight.quant_state, residual=x,
                                        allreduce=self.fused_ar, prefetch=self._hint(nxt))
                    continue
                g, u = gemv_4bit_fused(x, None, group=L.gate_up, rms_weight=L.ln2, rms_eps=cfg.eps,
                                       prefetch=self._hint(L.down_proj)).split(L.gate_up.splits, dim=-1)
                x = gemv_4bit_fused(u, L.down_proj.weight.data, L.down_proj.weight.quant_state, gate=g, residual=x,
                                    allreduce=self.fused_ar, prefetch=self._hint(nxt))
                continue
            x = x + self._allreduce(L.o_proj(a))
            h = F.rms_norm(x, (cfg.hidden,), L.ln2, cfg.eps)
            if L.gate_up is not None and T == 1:
                g, u = L.gate_up(h)
            else:
                g, u = L.gate_proj(h), L.up_proj(h)
            x = x + self._allreduce(L.down_proj(F.silu(g) * u))
        x = F.rms_norm(x[:, -1:], (cfg.hidden,), self.norm, cfg.eps)
        return F.linear(x, self.lm_head).view(-1)

    def _decode_chained(self, x, pos):
        """One decode step in 1 + 2 * layers launches: [norm + q/k/v of layer 0], then per layer the attention glue kernel and
        one chained launch [o + residual -> norm + gate/up -> SwiGLU + down + residual -> norm + q/k/v of the NEXT layer]."""
        cfg, Ls = self.cfg, self.layers
        h = x.clone()                                               # residual stream, updated in place by the GEMV epilogues
        I = Ls[0].gate_up.splits[0]
        qkv = gemv_4bit_fused(h, None, group=Ls[0].qkv, rms_weight=Ls[0].ln1, rms_eps=cfg.eps)
        gu = torch.empty(1, 1, Ls[0].gate_up.out_features, dtype=h.dtype, device=h.device)
        for li, L in enumerate(Ls):
            a = decode_attention(qkv, self.cos, self.sin, self.k_cache[li], self.v_cache[li], pos, L.nh, L.nkv, flags=self.attn_flags)
            with gemv_4bit_chain() as ch:
                ch.add(a, L.o_proj.weight.data, L.o_proj.weight.quant_state, residual=h, out=h, allreduce=self.fused_ar)
                ch.add(h, None, group=L.gate_up, rms_weight=L.ln2, rms_eps=cfg.eps, out=gu)
                ch.add(gu[..., I:], L.down_proj.weight.data, L.down_proj.weight.quant_state, gate=gu[..., :I], residual=h, out=h,
                       allreduce=self.fused_ar)
                if li + 1 < len(Ls):
                    N = Ls[li + 1]
                    ch.add(h, None, group=N.qkv, rms_weight=N.ln1, rms_eps=cfg.eps, out=qkv)
        y = F.rms_norm(h, (cfg.hidden,), self.norm, cfg.eps)
        return F.linear(y, self.lm_head).view(-1)

    @torch.no_grad()
    def generate(self, prompt: torch.Tensor, new_tokens: int, use_graph: bool = True):
        """Greedy decode: prefill `prompt` [P], then `new_tokens` decode steps.  Returns (tokens [new_tokens], decode seconds)."""
        import time

        dev = prompt.device
        P = prompt.shape[0]
        # positions written: P prompt tokens, `new_tokens` decode steps, plus the warm-up / capture steps of the graph path (their
        # tok / pos are restored, but they do run and do write cache rows P .. P+2)
        last = P + new_tokens + (3 if use_graph else 0)
        if last > self.cfg.max_len:
            raise ValueError(f"prompt ({P}) + new_tokens ({new_tokens})" + (" + 3 graph warm-up steps" if use_graph else "")
                             + f" = {last} exceeds the KV cache (cfg.max_len = {self.cfg.max_len})")
        logits = self.forward(prompt, torch.arange(P, device=dev))
        tok = logits.argmax().view(1)
        pos = torch.full((1,), P, device=dev, dtype=torch.int64)
        out = torch.empty(new_tokens, dtype=torch.int64, device=dev)

        def step():
            lg = self.forward(tok, pos)
            argmax(lg, out=tok)  # greedy: one short launch straight into the next step's input token
            pos.add_(1)

        graph = None
        if use_graph:
            from .graphs import capture

            tok0, pos0 = tok.clone(), pos.clone()
            graph = capture(step, warmup=2)           # warm-up + capture advance tok/pos: restore them
            tok.copy_(tok0)
            pos.copy_(pos0)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(new_tokens):
            out[i : i + 1].copy_(tok)
            if graph is not None:
                graph.replay()
            else:
                step()
        torch.cuda.synchronize(dev)
        return out, time.perf_counter() - t0


def linear4bit_factory(device, dtype=torch.bfloat16, quant_type="nf4", seed=0, tp_rank=0, tp=1, col_names=("q_proj", "k_proj", "v_proj", "gate_proj", "up_proj")):
    """Random-init Linear4bit projections (weights ~ N(0, 0.02^2), HF initializer_range), quantised on creation.
    Under TP every rank draws the FULL weight from the shared seed and quantises its own slice."""
    from .core import Params4bit
    from .modules import Linear4bit

    counter = [seed]

    def make(fin, fout, name):
        counter[0] += 1
        g = torch.Generator(device=device).manual_seed(counter[0])
        col = name.split(".")[-1] in col_names
        full_out, full_in = (fout * tp, fin) if col else (fout, fin * tp)
        W = (torch.randn(full_out, full_in, device=device, dtype=torch.float32, generator=g) * 0.02).to(dtype)
        if tp > 1:
            W = (W[tp_rank * fout:(tp_rank + 1) * fout] if col else W[:, tp_rank * fin:(tp_rank + 1) * fin]).contiguous()
        lin = Linear4bit(fin, fout, bias=False, compute_dtype=dtype, quant_type=quant_type, device="meta")
        lin.weight = Params4bit(W, requires_grad=False, quant_type=quant_type, module=lin).to(device)
        return lin

    return make


def dense_factory(device, dtype=torch.bfloat16, seed=0):
    """Unquantised nn.Linear projections from the same seeds (the "HF native" baseline of the reference's README)."""
    counter = [seed]

    def make(fin, fout, name):
        counter[0] += 1
        g = torch.Generator(device=device).manual_seed(counter[0])
        lin = nn.Linear(fin, fout, bias=False, device=device, dtype=dtype)
        lin.weight.data = (torch.randn(fout, fin, device=device, dtype=torch.float32, generator=g) * 0.02).to(dtype)
        return lin

    return make
