#!/bin/bash
# round 2, call G (re-entry): re-establish the baselines -- ring kernel timings and traces, bench line, full GPU test suite
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
T=gpurun_out/r2g_timing.log
: > $T
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2g_smi.txt 2>&1
for cfg in "16 2" "8 1"; do
  set -- $cfg
  timeout 120 $B --nc $1 --wps $2 --chain 4 2>&1 | grep -E "RESULT|FAIL|error" >> $T
done
timeout 120 $B --nc 16 --wps 2 --chain 1 2>&1 | grep -E "RESULT|FAIL|error" >> $T
echo "nocompute:" >> $T
timeout 120 ${B}_nocompute --nc 16 --wps 2 --chain 4 2>&1 | grep -E "RESULT|error" >> $T
timeout 120 ${B}_nocompute --nc 8 --wps 1 --chain 4 2>&1 | grep -E "RESULT|error" >> $T
timeout 120 $B --nc 16 --wps 2 --chain 4 --trace > gpurun_out/r2g_trace_16_2.log 2>&1
timeout 120 ${B}_nocompute --nc 16 --wps 2 --chain 4 --trace > gpurun_out/r2g_trace_nocompute.log 2>&1
cat $T
grep -A40 "launch 16" gpurun_out/r2g_trace_16_2.log | head -44
echo "== nocompute trace"
grep -A40 "launch 16" gpurun_out/r2g_trace_nocompute.log | head -44
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -c 1500 gpurun_out/r2g_bench.json
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest_all.log 2>&1
tail -5 gpurun_out/r2g_pytest_all.log
