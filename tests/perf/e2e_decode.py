"""End-to-end batch-1 decode tok/s (BASELINE config 3): random-init Llama-3-8B, 32-token prompt, 60 new tokens, greedy.

    python tests/perf/e2e_decode.py [--layers 32] [--which ours-graph,ours-eager,ref-asshipped,ref-bf16,dense]
"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quantizations_b200 import llama

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--new", type=int, default=60)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--which", default="ours-graph,ours-eager,ref-bf16,ref-asshipped,dense")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = llama.LlamaConfig(layers=a.layers)
prompt = torch.arange(1, 33, device=dev)
res = {}
for which in a.which.split(","):
    torch.cuda.empty_cache()
    if which.startswith("ours"):
        make, dtype = llama.linear4bit_factory(dev, torch.bfloat16, "nf4"), torch.bfloat16
    elif which == "dense":
        make, dtype = llama.dense_factory(dev, torch.bfloat16), torch.bfloat16
    else:
        from oracle import ref_linear
        if not ref_linear.available():
            print(which, "unavailable (oracle/_ref missing)"); continue
        if which == "ref-asshipped":   # fp16 model, FP4, fp32 compute: the reference's README configuration
            make, dtype = ref_linear.ref_factory(dev, torch.float16, "fp4", None), torch.float16
        else:                          # same config as ours: bf16 model, NF4 table, bf16 GEMV instance
            make, dtype = ref_linear.ref_factory(dev, torch.bfloat16, "nf4", torch.bfloat16), torch.bfloat16
    model = llama.Llama(cfg, make, dev, dtype)
    use_graph = which == "ours-graph" or which == "dense"
    ctx = torch.cuda.stream(torch.cuda.default_stream(dev)) if which.startswith("ref") else torch.cuda.stream(torch.cuda.Stream())
    with ctx:
        toks, _ = model.generate(prompt, 8, use_graph=False)   # warm-up (also: first-call probes, cuBLAS handles)
        best = None
        for _ in range(a.iters):
            toks, dt = model.generate(prompt, a.new, use_graph=use_graph)
            best = dt if best is None else min(best, dt)
    res[which] = {"tok_s": round(a.new / best, 1), "ms_per_token": round(best / a.new * 1e3, 3), "first_tokens": toks[:6].tolist()}
    print(which, res[which], flush=True)
    del model
print(json.dumps(res))
