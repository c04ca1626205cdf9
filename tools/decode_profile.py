import sys, os, torch
sys.path.insert(0, "/root/repo")
from quantizations_b200 import llama
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
cfg = llama.LlamaConfig(layers=int(os.environ.get("LAYERS", "32")))
m = llama.Llama(cfg, llama.linear4bit_factory(dev, torch.bfloat16, "nf4"), dev, torch.bfloat16)
prompt = torch.arange(1, 33, device=dev)
toks, dt = m.generate(prompt, 4, use_graph=False)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
tok = torch.tensor([5], device=dev); pos = torch.tensor([40], device=dev)
for _ in range(3): m.forward(tok, pos)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        lg = m.forward(tok, pos); t = lg.argmax()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
