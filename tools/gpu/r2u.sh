#!/bin/bash
# round 2, call U: full GPU suite + bench after the small-batch kernel went onto the default path
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest_all.log 2>&1
tail -8 gpurun_out/r2u_pytest_all.log
timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
tail -c 600 gpurun_out/r2u_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d.get('ring_check'))
for k,v in d['gemv_sweep']['shapes'].items(): print(k, v)
P
