#!/bin/bash
# round 2, final: what the driver runs -- GPU test suite, smoke, both bench arms with default flags
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=$SECONDS
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2fin_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2fin_pytest_gpu.log; echo "pytest $((SECONDS-S)) s"; S=$SECONDS
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/r2fin_bench_reference.json 2> gpurun_out/r2fin_bench_reference.err
echo "reference arm $((SECONDS-S)) s rc=$?"; S=$SECONDS
timeout 900 python bench.py > gpurun_out/r2fin_bench_ours.json 2> gpurun_out/r2fin_bench_ours.err
echo "our arm $((SECONDS-S)) s rc=$?"
tail -c 300 gpurun_out/r2fin_bench_ours.err
python - <<'P'
import json
o=json.loads(open('gpurun_out/r2fin_bench_ours.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2fin_bench_reference.json').read().strip().splitlines()[-1])
print("ours", o['value'], o['ms_per_step'], o['roofline']['frac'], o['roofline']['traffic'], "e2e", o['e2e']['value'], o['gpu_launches'], o['clocks'])
print("ref", r['value'], r['ms_per_step'], "e2e", r['e2e']['value'], "ratio e2e", o['e2e']['value']/r['e2e']['value'])
print(o.get('ring_check'), o['decode'], o.get('tp70b',{}).get('ms_per_step'))
for k,v in o['gemv_sweep']['shapes'].items(): print(k, v['us'], v['frac'], v['tokens4_us'], v['tokens8_us'])
print(o['blockwise']['quantize']['frac'], o['blockwise']['dequantize']['frac'])
P
