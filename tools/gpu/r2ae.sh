#!/bin/bash
# round 2, call AE: final verification -- smoke, full GPU suite, default bench, reference arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ae_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2ae_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2ae_pytest.log 2>&1; tail -3 gpurun_out/r2ae_pytest.log
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench rc=$? wall $SECONDS s"
timeout 600 python bench.py --impl reference > gpurun_out/r2ae_bench_ref.json 2>> gpurun_out/r2ae_bench.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2ae_bench.json")); r = json.load(open("gpurun_out/r2ae_bench_ref.json"))
print("ours", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "decode", d["decode"]["tok_s"], "clocks", d["clocks"])
print("ref ", r["value"], r["ms_per_step"], "same workload:", d["config"]["workload"] == r["config"]["workload"])
print("ratio e2e", d["e2e"]["value"] / r["e2e"]["value"], "ratio value", d["value"] / r["value"])
PY
