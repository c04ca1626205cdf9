"""BASELINE config 3 with the reference README's own protocol (README.md:104-126): a random-init Llama-3-8B built by HF transformers,
`model.generate(max_new_tokens=60, min_new_tokens=60, do_sample=False, use_cache=True)`, batch 1, 1 warm-up + 5 timed iterations,
TPS = 60 / mean time -- with the Linear layers (a) left dense (HF native, bf16), (b) swapped for this engine's Linear4bit through
transformers' own replace_with_bnb_linear (quantizations_b200.hf), (c) swapped for the reference's kernels (oracle/ref_linear; FP4, the
reference's only quantiser).  Lives under tests/ because it touches oracle/ (test infrastructure).  No tokenizer / checkpoint offline: input_ids = arange(1, 33).

    python tests/perf/hf_generate_tps.py [--layers 32] [--which native,ours,reference]

HF generate() is host-bound at batch 1 (hundreds of small launches and Python per token), so this measures the engine inside that
harness; the CUDA-graph decode loop of quantizations_b200.llama (bench.py's `decode` leg) is the engine without it.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def build(layers, dev):
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=layers, num_attention_heads=32, num_key_value_heads=8,
                      vocab_size=128256, rope_theta=500000.0, rms_norm_eps=1e-5, max_position_embeddings=8192, tie_word_embeddings=False)
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(old)
    return model.eval()


def swap_reference(model, dev):
    import quantizations_b200 as q
    from oracle import ref_linear

    for name, m in list(model.named_modules()):
        for cname, child in list(m.named_children()):
            if type(child) is torch.nn.Linear and cname != "lm_head":
                packed, st = q.quantize_4bit(child.weight.detach().to(dev), quant_type="fp4")  # the reference's only quantiser
                setattr(m, cname, ref_linear.RefLinear4bit(packed, st, torch.bfloat16))
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--which", default="native,ours,reference,ours-graph,native-graph")
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    ids = torch.arange(1, 33, device=dev).view(1, -1)
    res = {}
    for which in a.which.split(","):
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        # "ours" is built on the host and quantised layer by layer on its way to the GPU (what from_pretrained(load_in_4bit=True) does):
        # the dense model never exists in device memory
        model = build(a.layers, "cpu" if which == "ours" else dev)
        if which.startswith("ours"):
            from transformers import BitsAndBytesConfig

            from quantizations_b200 import hf

            hf.quantize_model(model, BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type="nf4", bnb_4bit_use_double_quant=True,
                                                        bnb_4bit_compute_dtype=torch.bfloat16), device=dev)
        elif which == "reference":
            from oracle import ref_linear

            if not ref_linear.available():
                print("reference unavailable (oracle/_ref missing)")
                continue
            swap_reference(model, dev)
        if which.endswith("-graph"):  # the same HF model, single-token step replayed as a CUDA graph (quantizations_b200.hf.graph_generate)
            from quantizations_b200 import hf as q_hf

            with torch.cuda.stream(torch.cuda.Stream()):
                toks, _ = q_hf.graph_generate(model, ids, 8)
                ts = [q_hf.graph_generate(model, ids, 60)[1] for _ in range(3)]
                eager, te = q_hf.graph_generate(model, ids, 60, use_graph=False)
                toks, _ = q_hf.graph_generate(model, ids, 60)
            res[which] = {"tps": round(60 / (sum(ts) / len(ts)), 1), "best_tps": round(60 / min(ts), 1),
                          "same_loop_eager_tps": round(60 / te, 1), "tokens_equal_to_eager_loop": int((toks == eager).sum().item())}
            print(which, res[which], flush=True)
            del model
            continue
        kw = dict(max_new_tokens=60, min_new_tokens=60, do_sample=False, use_cache=True, pad_token_id=0)
        load_peak = torch.cuda.max_memory_allocated()
        with torch.no_grad():
            out = model.generate(ids, **kw)  # warm-up
            torch.cuda.synchronize()
            ts = []
            for _ in range(a.iters):
                t0 = time.perf_counter()
                out = model.generate(ids, **kw)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
        assert out.shape == (1, 92)
        res[which] = {"tps": round(60 / (sum(ts) / len(ts)), 1), "best_tps": round(60 / min(ts), 1), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2),
                      "load_peak_mem_GB": round(load_peak / 1e9, 2)}
        print(which, res[which], flush=True)
        del model
        torch.cuda.reset_peak_memory_stats()
    print(json.dumps({"protocol": "HF generate, random-init Llama-3-8B (%d layers), bs=1, 32-token prompt, 60 new tokens, greedy, 1 warm-up + %d iterations" % (a.layers, a.iters),
                      "transformers": __import__("transformers").__version__, **res}))


if __name__ == "__main__":
    main()
