#!/bin/bash
# round 2, call Y2 (2 GPUs): tensor-parallel tests and bench after the ring kernel's 32-stage change
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tp_gpu.py -x -q > gpurun_out/r2y2_pytest_tp.log 2>&1
tail -4 gpurun_out/r2y2_pytest_tp.log
for mode in "" "--ring-tp"; do
  tag=${mode:+_ring}
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-tp70b --no-sweep $mode > gpurun_out/r2y2_bench_tp2$tag.json 2> gpurun_out/r2y2_bench_tp2$tag.err
  tail -2 gpurun_out/r2y2_bench_tp2$tag.err | cut -c1-200
  python - "$tag" <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/r2y2_bench_tp2%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], d["ms_per_step"], d["value"], d.get("tp_check"), d.get("ring_check"), d["gpu_launches"])
P
done
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2y2_bench_tp2.json") if l.startswith("{")][-1])
print(json.dumps(d.get("allreduce"), indent=1))
P
