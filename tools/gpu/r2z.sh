#!/bin/bash
# round 2, call Z (8 GPUs): the 8B stack tensor-parallel over 8 ranks, single launches vs ring launches; TP test on 8 ranks is not needed (2-rank test covers the protocol)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for mode in "" "--ring-tp"; do
  tag=$([ -z "$mode" ] && echo single || echo ring)
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-tp70b $mode > gpurun_out/r2z_bench_tp8_$tag.json 2> gpurun_out/r2z_bench_tp8_$tag.err
  echo "$tag rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2z_bench_tp8_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("tp_check", {}).get("status"), d.get("ring_check"), "decode", d.get("decode", {}).get("tok_s"))
except Exception as e:
    print("$tag failed", e)
PY
done
tail -5 gpurun_out/r2z_bench_tp8_ring.err | cut -c1-200
