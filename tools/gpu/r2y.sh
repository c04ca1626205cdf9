#!/bin/bash
# round 2, call Y (2 GPUs): tensor-parallel tests + bench at N = 2 with the dependent data flow and the tp70b leg
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_tp_gpu.py -x -q -m gpu -rs > gpurun_out/r2y_pytest_tp.log 2>&1
tail -6 gpurun_out/r2y_pytest_tp.log
SECONDS=0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2y_bench_tp2.json 2> gpurun_out/r2y_bench_tp2.err
echo "tp2 bench wall seconds: $SECONDS rc=$?"
tail -3 gpurun_out/r2y_bench_tp2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2y_bench_tp2.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "tp_check", d.get("tp_check"))
print("decode", d.get("decode"))
print("tp70b", d.get("tp70b"))
PY
