#!/bin/bash
# round 2, call S: up to 32 stages per ring launch -- micro-benchmark at 8 / 16 / 32 stages per launch, ring tests, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
T=gpurun_out/r2s_timing.log
: > $T
for ch in 8 16 32; do
  timeout 120 $B --nc 16 --wps 2 --chain $ch 2>&1 | grep -E "RESULT|FAIL|error|mismatch" >> $T
done
cat $T
timeout 900 python -m pytest tests/test_ring_gpu.py tests/test_gpu_parity.py -x -q > gpurun_out/r2s_pytest.log 2>&1
tail -4 gpurun_out/r2s_pytest.log
timeout 600 python bench.py --no-sweep --no-tp70b > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
tail -c 400 gpurun_out/r2s_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d.get('ring_check'), d['gpu_launches'], d['e2e']['ms_per_step'], d['decode'])
print(d['launch'])
P
