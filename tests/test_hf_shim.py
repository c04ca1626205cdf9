"""HF integration shim (quantizations_b200/hf.py): transformers' own replace_with_bnb_linear builds this
engine's Linear4bit; a tiny random-init HF Llama quantised through it decodes like its dense twin with dequantised weights."""
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _tiny():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=512, max_position_embeddings=128, tie_word_embeddings=False)
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg)


def test_hf_replace_function_builds_our_linear4bit():
    from transformers import BitsAndBytesConfig

    from quantizations_b200 import hf
    import quantizations_b200 as q

    ns = hf.install()
    import bitsandbytes as bnb  # resolves to the namespace (no real bitsandbytes in this image)

    assert bnb.nn.Linear4bit is q.Linear4bit and bnb.functional.dequantize_4bit is q.dequantize_4bit and bnb is ns
    model = _tiny()
    cfg = BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type="nf4", bnb_4bit_use_double_quant=True,
                             bnb_4bit_compute_dtype=torch.bfloat16)
    hf.replace_with_bnb_linear(model, cfg, modules_to_not_convert=["lm_head"])
    swapped = [n for n, m in model.named_modules() if isinstance(m, q.Linear4bit)]
    assert len(swapped) == 2 * 7 and all("lm_head" not in n for n in swapped)
    lin = model.model.layers[0].self_attn.q_proj
    assert lin.weight.device.type == "meta" and lin.weight.quant_type == "nf4" and lin.weight.compress_statistics
    assert lin.compute_dtype == torch.bfloat16 and lin.source_cls is torch.nn.Linear
    assert type(model.lm_head) is torch.nn.Linear


@pytest.mark.gpu
def test_hf_llama_quantised_through_the_shim_matches_dense_twin():
    from transformers import BitsAndBytesConfig

    from quantizations_b200 import hf
    import quantizations_b200 as q

    dev = "cuda:0"
    model = _tiny().to(torch.bfloat16)
    twin = _tiny().to(torch.bfloat16)
    twin.load_state_dict(model.state_dict())
    cfg = BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type="nf4", bnb_4bit_use_double_quant=True,
                             bnb_4bit_compute_dtype=torch.bfloat16)
    hf.quantize_model(model, cfg, device=dev)
    twin.to(dev)
    for (n, m), (_, t) in zip(model.named_modules(), twin.named_modules()):
        if isinstance(m, q.Linear4bit):
            assert m.weight.dtype == torch.uint8 and m.weight.quant_state is not None
            t.weight.data = q.dequantize_4bit(m.weight.data, m.weight.quant_state).t().contiguous().to(torch.bfloat16)
    ids = torch.arange(3, 19, device=dev).view(1, -1)
    with torch.no_grad():
        lq, ld = model(ids).logits.float(), twin(ids).logits.float()              # prefill: fused GEMM / dequant + cuBLAS
        assert (lq - ld).abs().max().item() <= 3e-2 * ld.abs().max().item()
        one = ids[:, :1]
        sq, sd = model(one).logits.float(), twin(one).logits.float()              # single token: the decode GEMV
        assert (sq - sd).abs().max().item() <= 3e-2 * sd.abs().max().item()
        out = model.generate(ids, max_new_tokens=8, do_sample=False)
        assert out.shape == (1, 24)


@pytest.mark.gpu
def test_graph_generate_replays_the_eager_static_cache_loop():
    """quantizations_b200.hf.graph_generate: the HF model's single-token forward captured in a CUDA graph yields exactly the
    tokens of the same static-cache loop run eagerly, and the first token is generate()'s (later ones may differ from the dynamic
    cache path by rounding on a random-init model)."""
    from transformers import BitsAndBytesConfig

    from quantizations_b200 import hf

    dev = "cuda:0"
    model = _tiny().to(torch.bfloat16)
    cfg = BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type="nf4", bnb_4bit_use_double_quant=True,
                             bnb_4bit_compute_dtype=torch.bfloat16)
    hf.quantize_model(model, cfg, device=dev)
    ids = torch.arange(3, 19, device=dev).view(1, -1)
    with torch.cuda.stream(torch.cuda.Stream()):
        eager, _ = hf.graph_generate(model, ids, 12, use_graph=False)
        graph, _ = hf.graph_generate(model, ids, 12)
        ref = model.generate(ids, max_new_tokens=1, do_sample=False, pad_token_id=0)
    assert torch.equal(eager, graph)
    assert int(ref[0, -1]) == int(graph[0, 0])


@pytest.mark.gpu
def test_transformers_own_conversion_ops_build_and_reload_our_weights():
    """What from_pretrained(quantization_config=BitsAndBytesConfig(load_in_4bit=True)) does per weight, without `accelerate` (absent in
    this image, so from_pretrained itself cannot run): transformers' OWN conversion ops -- integrations/bitsandbytes.py
    `Bnb4bitQuantize.convert` (dense checkpoint tensor -> bnb.nn.Params4bit(value, **old.__dict__).to(device)) and
    `Bnb4bitDeserialize.convert` (pre-quantised checkpoint: packed weight + the bnb statistics keys ->
    bnb.nn.Params4bit.from_prequantized) -- executed unmodified with `bitsandbytes` resolving to this engine.  The quantised model
    must decode like its dense twin, and a model rebuilt from the state dict of the first (the serialised bnb key set) must
    reproduce it bit for bit.  Reference: core.py:91-190 (Params4bit), modules.py:67-151."""
    from transformers import BitsAndBytesConfig
    from transformers.integrations.bitsandbytes import Bnb4bitDeserialize, Bnb4bitQuantize

    from quantizations_b200 import hf
    import quantizations_b200 as q

    dev = "cuda:0"
    cfg = BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type="nf4", bnb_4bit_use_double_quant=True,
                             bnb_4bit_compute_dtype=torch.bfloat16)
    dense = _tiny().to(torch.bfloat16)
    checkpoint = {k: v.clone() for k, v in dense.state_dict().items()}

    def skeleton():
        m = _tiny().to(torch.bfloat16)
        hf.replace_with_bnb_linear(m, cfg, modules_to_not_convert=["lm_head"])
        return m

    # ---- dense checkpoint -> quantise on load
    model = skeleton()
    op = Bnb4bitQuantize(None)
    names = [n for n, m in model.named_modules() if isinstance(m, q.Linear4bit)]
    assert len(names) == 14  # 2 layers x (q, k, v, o, gate, up, down)
    for n in names:
        out = op.convert({n + ".weight": [checkpoint[n + ".weight"].to(dev)]}, full_layer_name=n + ".weight", model=model)
        (key, value), = out.items()
        assert key == n + ".weight" and isinstance(value, q.Params4bit) and value.dtype == torch.uint8 and value.quant_state is not None
        mod = model.get_submodule(n)
        mod.weight = value
    rest = {k: v for k, v in checkpoint.items() if not any(k == n + ".weight" for n in names)}
    missing, unexpected = model.load_state_dict(rest, strict=False)
    assert not unexpected and all(any(k.startswith(n + ".weight") for n in names) for k in missing)
    model.to(dev)
    twin = dense.to(dev)
    for n in names:
        m = model.get_submodule(n)
        twin.get_submodule(n).weight.data = q.dequantize_4bit(m.weight.data, m.weight.quant_state).t().contiguous().to(torch.bfloat16)
    ids = torch.arange(3, 19, device=dev).view(1, -1)
    with torch.no_grad():
        lq, ld = model(ids[:, :1]).logits.float(), twin(ids[:, :1]).logits.float()
        assert (lq - ld).abs().max().item() <= 3e-2 * ld.abs().max().item()

    # ---- pre-quantised checkpoint (the state dict of the model above: bnb's key names) -> deserialise on load
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    assert any(k.endswith("weight.quant_state.bitsandbytes__nf4") for k in sd) and any(k.endswith("weight.nested_absmax") for k in sd)
    again = skeleton()
    dop = Bnb4bitDeserialize(None)
    for n in names:
        pre = n + ".weight"
        inp = {"weight": [sd[pre].to(dev)]}
        inp.update({k[len(n) + 1:]: [v.to(dev)] for k, v in sd.items() if k.startswith(pre + ".")})
        out = dop.convert(inp, model=again, full_layer_name=pre)
        value = out["weight"]
        assert isinstance(value, q.Params4bit) and value.quant_state is not None
        again.get_submodule(n).weight = value
    again.load_state_dict({k: v for k, v in sd.items() if not any(k.startswith(n + ".weight") for n in names)}, strict=False)
    again.to(dev)
    with torch.no_grad():
        assert torch.equal(again(ids[:, :1]).logits, model(ids[:, :1]).logits)
        assert torch.equal(again(ids).logits, model(ids).logits)
