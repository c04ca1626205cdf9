#!/bin/bash
# round 2, call Z2 (8 GPUs): the driver's scaling command at N = 8 with the final code (tp_check, allreduce latency, tp70b leg)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2z2_bench_tp8.json 2> gpurun_out/r2z2_bench_tp8.err
echo "rc=$? wall $((SECONDS-S)) s"
tail -3 gpurun_out/r2z2_bench_tp8.err | cut -c1-300
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2z2_bench_tp8.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["value"], d.get("tp_check"), d.get("decode"))
print(json.dumps(d.get("allreduce")))
print({k: (v if not isinstance(v, dict) else {kk: v[kk] for kk in list(v)[:3]}) for k, v in (d.get("tp70b") or {}).items()})
P
