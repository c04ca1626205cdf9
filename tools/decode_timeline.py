"""Developer tool: kernel timeline of ONE replay of the captured decode step (start offset, duration, gap to the previous kernel)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantizations_b200 import llama, graphs
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
cfg = llama.LlamaConfig(layers=int(os.environ.get("LAYERS", "32")))
m = llama.Llama(cfg, llama.linear4bit_factory(dev, torch.bfloat16, "nf4"), dev, torch.bfloat16)
m.generate(torch.arange(1, 33, device=dev), 4, use_graph=False)
tok = torch.tensor([5], device=dev); pos = torch.tensor([40], device=dev)
def step():
    lg = m.forward(tok, pos); tok.copy_(lg.argmax().view(1))
g = graphs.capture(step)
for _ in range(3): g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay(); torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
tot_busy = 0; prev_end = t0; rows = []
for e in ev:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    rows.append((s, d, e.time_range.start - prev_end, e.name[:60])); tot_busy += d; prev_end = max(prev_end, e.time_range.end)
print(f"kernels {len(ev)}  span {prev_end - t0:.1f} us  sum of durations {tot_busy:.1f} us")
import collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for s, d, gap, n in rows:
    agg[n][0] += 1; agg[n][1] += d; agg[n][2] += max(gap, 0)
for n, (c, d, gp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c:4d} x {n:60s} dur {d:8.1f} us  gaps before {gp:8.1f} us")
for r in rows[:14]: print("  %8.1f us  dur %6.1f  gap %5.1f  %s" % r)
