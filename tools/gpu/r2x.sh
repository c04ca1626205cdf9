#!/bin/bash
# round 2, call X: full GPU suite + the default bench line (ring kernel, gemv_sweep, tp70b leg)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
true
true
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench wall seconds: $SECONDS"

tail -3 gpurun_out/r2x_bench.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2x_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("decode", d.get("decode"))
print("sweep", {k: (v["us"], v["frac"]) for k, v in d["gemv_sweep"]["shapes"].items()})
print("tp70b", d.get("tp70b"))
print("cpu", d.get("cpu_baseline", {}).get("value"), "blockwise", {k: v["frac"] for k, v in d["blockwise"].items() if isinstance(v, dict)})
PY

import json; d=json.load(open('gpurun_out/r2x_bench_ref.json')); print('ref', d['value'], d['ms_per_step'], d['config'])"
