"""A `bitsandbytes`-shaped namespace over this engine, and the glue that lets HF transformers' own bnb integration use it
(the reference asks its users to patch transformers by hand instead, README.md:52-86).

transformers' 4-bit path touches exactly these names (transformers 5.5 integrations/bitsandbytes.py):
    bnb.nn.Linear4bit(in, out, bias, compute_dtype, compress_statistics=, quant_type=, quant_storage=)     :203-211
    bnb.nn.Params4bit(value, requires_grad=False, **old.__dict__).to(device)                              :56
    bnb.nn.Params4bit.from_prequantized(data=, quantized_stats=, requires_grad=, device=, module=)        :85-91
    bnb.functional.dequantize_4bit(weight.data, weight.quant_state)                                       :249
`namespace()` builds a module object with those attributes; `install()` registers it as `bitsandbytes` (only if no real
bitsandbytes is importable) and binds it inside transformers' integration module; `replace_with_bnb_linear` then runs
transformers' OWN replacement function, unmodified, against this engine.  `from_pretrained(load_in_4bit=True)` itself also
insists on the `accelerate` package and on bitsandbytes' installed-distribution metadata (quantizers/quantizer_bnb_4bit.py:
56-64); neither exists in this image, so `quantize_model` performs the two steps from_pretrained would (replace, then
requantise every weight the way Bnb4bitQuantize.convert does) on an already constructed model.
"""
from __future__ import annotations

import sys
import types

import torch

from . import core, modules

__all__ = ["namespace", "install", "replace_with_bnb_linear", "quantize_model", "graph_generate"]

_NS = None


def namespace() -> types.ModuleType:
    """The bitsandbytes-shaped module (bnb.nn.*, bnb.functional.*) backed by quantizations_b200."""
    global _NS
    if _NS is None:
        bnb = types.ModuleType("bitsandbytes")
        bnb.__version__ = "0.46.0+quantizations_b200"
        bnb.__doc__ = "bitsandbytes-compatible namespace provided by quantizations_b200 (4-bit Linear path only)"
        nn_ = types.ModuleType("bitsandbytes.nn")
        nn_.Linear4bit = modules.Linear4bit
        nn_.Params4bit = core.Params4bit
        fn = types.ModuleType("bitsandbytes.functional")
        for name in ("quantize_4bit", "dequantize_4bit", "gemv_4bit", "quantize_blockwise", "dequantize_blockwise", "QuantState",
                     "get_4bit_type", "create_dynamic_map"):
            setattr(fn, name, getattr(core, name))
        bnb.nn, bnb.functional = nn_, fn
        bnb.matmul_4bit = modules.matmul_4bit
        import importlib.machinery

        for mod in (bnb, nn_, fn):  # importlib.util.find_spec() on a module already in sys.modules requires a spec
            mod.__spec__ = importlib.machinery.ModuleSpec(mod.__name__, loader=None)
        _NS = bnb
    return _NS


def install(force: bool = False) -> types.ModuleType:
    """Make `import bitsandbytes` resolve to this engine (unless a real bitsandbytes is installed and `force` is False) and bind
    it as `bnb` inside transformers.integrations.bitsandbytes, where the replacement / conversion code looks it up."""
    ns = namespace()
    have_real = False
    if not force and "bitsandbytes" not in sys.modules:
        import importlib.util

        have_real = importlib.util.find_spec("bitsandbytes") is not None
    if force or not have_real:
        sys.modules["bitsandbytes"] = ns
        sys.modules["bitsandbytes.nn"] = ns.nn
        sys.modules["bitsandbytes.functional"] = ns.functional
    try:
        import transformers.integrations.bitsandbytes as hf_bnb

        hf_bnb.bnb = ns
    except ImportError:  # transformers absent: the namespace is still usable on its own
        pass
    return ns


def replace_with_bnb_linear(model: torch.nn.Module, quantization_config, modules_to_not_convert=None, pre_quantized: bool = False):
    """transformers.integrations.bitsandbytes.replace_with_bnb_linear, unmodified, constructing THIS engine's Linear4bit."""
    install()
    import transformers.integrations.bitsandbytes as hf_bnb

    return hf_bnb.replace_with_bnb_linear(model, modules_to_not_convert=modules_to_not_convert,
                                          quantization_config=quantization_config, pre_quantized=pre_quantized)


def quantize_model(model: torch.nn.Module, quantization_config, device="cuda", modules_to_not_convert=("lm_head",)):
    """What from_pretrained(quantization_config=BitsAndBytesConfig(load_in_4bit=True)) does to an already built model: swap every
    nn.Linear (except `modules_to_not_convert`) for Linear4bit through transformers' own replace function, then rebuild each
    weight exactly as Bnb4bitQuantize.convert does (integrations/bitsandbytes.py:56) -- which quantises it on the move to CUDA."""
    dense = {name: (m.weight.detach(), None if m.bias is None else m.bias.detach())
             for name, m in model.named_modules() if type(m) is torch.nn.Linear}
    replace_with_bnb_linear(model, quantization_config, modules_to_not_convert=list(modules_to_not_convert))
    ns = namespace()
    for name, m in list(model.named_modules()):
        if isinstance(m, ns.nn.Linear4bit):
            # one layer at a time, and the dense weight is released as soon as it is packed (the replaced nn.Linear modules are gone:
            # `dense` holds the last reference) -- the peak stays at the dense model's own footprint and falls from there, instead of
            # dense + quantised (19.7 vs 16.1 GB on Llama-3-8B, profiles/r01c_hf_generate.json)
            w, b = dense.pop(name)
            old = m.weight
            m.weight = ns.nn.Params4bit(w.to(device), requires_grad=False, **old.__dict__).to(device)
            if b is not None:
                m.bias = torch.nn.Parameter(b.to(device), requires_grad=False)
            del w, b, old
    dense.clear()
    return model.to(device)


@torch.no_grad()
def graph_generate(model: torch.nn.Module, input_ids: torch.Tensor, max_new_tokens: int, max_cache_len: int = None,
                   use_graph: bool = True):
    """Greedy batch-1 decoding of an HF causal LM with transformers' StaticCache and the single-token forward captured ONCE in a
    CUDA graph (quantizations_b200.graphs.capture) -- what `generate()` cannot do for the reference, whose kernels run on the
    legacy default stream.  HF's eager `generate()` spends ~15 ms of host time per token at batch 1 whatever
    the Linear layers cost; replayed as a graph the step costs what its kernels cost.  Works for any model whose forward accepts
    `past_key_values=StaticCache, cache_position=`; every Linear4bit launch of this engine is capturable.

    `use_graph=False` runs the same static-cache loop eagerly (the same kernels launch by launch: identical tokens).
    Returns (new tokens [1, max_new_tokens], seconds spent in the decode loop)."""
    import time

    from transformers import StaticCache

    from .graphs import capture

    if input_ids.shape[0] != 1:
        raise ValueError("graph_generate decodes one sequence")
    dev = input_ids.device
    P = input_ids.shape[1]
    if max_cache_len is not None and P + max_new_tokens + (3 if use_graph else 0) > max_cache_len:
        raise ValueError(f"prompt ({P}) + max_new_tokens ({max_new_tokens}) + graph warm-up steps exceed max_cache_len ({max_cache_len})")
    cache = StaticCache(config=model.config, max_cache_len=max_cache_len or P + max_new_tokens + 8)
    out = model(input_ids, past_key_values=cache, cache_position=torch.arange(P, device=dev), use_cache=True)
    tok = out.logits[:, -1].argmax(-1, keepdim=True)
    pos = torch.full((1,), P, device=dev, dtype=torch.long)
    new = torch.empty(1, max_new_tokens, dtype=torch.long, device=dev)

    def step():
        o = model(tok, past_key_values=cache, cache_position=pos, use_cache=True)
        lg = o.logits[:, -1]
        if lg.is_cuda and lg.is_contiguous() and lg.dtype in core._DTYPE_CODE and (lg.data_ptr() & 15) == 0:
            core.argmax(lg.view(-1), out=tok.view(-1))  # one ~3 us launch instead of torch's 40-us reduction over the vocabulary
        else:
            tok.copy_(lg.argmax(-1, keepdim=True))
        pos.add_(1)

    graph = None
    if use_graph:
        tok0, pos0 = tok.clone(), pos.clone()
        graph = capture(step, warmup=2)  # warm-up steps write cache rows >= P: masked until the real steps overwrite them
        tok.copy_(tok0)
        pos.copy_(pos0)
        for layer in getattr(cache, "layers", ()):  # transformers >= 5: a static layer counts its own write position on the device
            if isinstance(getattr(layer, "cumulative_length", None), torch.Tensor):
                layer.cumulative_length.fill_(P)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(max_new_tokens):
        new[:, i:i + 1].copy_(tok)
        if graph is not None:
            graph.replay()
        else:
            step()
    torch.cuda.synchronize(dev)
    return new, time.perf_counter() - t0
