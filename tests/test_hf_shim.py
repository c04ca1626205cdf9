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
