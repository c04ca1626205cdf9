#!/bin/bash
# round 2, call A: new parity tests + first ring-kernel measurements
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 120 $B --check-only --nc 16 > gpurun_out/r2a_check.log 2>&1; echo "check rc=$?" >> gpurun_out/r2a_check.log
tail -25 gpurun_out/r2a_check.log
if grep -q FAIL gpurun_out/r2a_check.log || ! grep -q "check rc=0" gpurun_out/r2a_check.log; then
  echo "ring check failed: skipping timing"
else
  for nc in 12 16 20 24; do
    for ch in 4 1; do
      timeout 120 $B --nc $nc --chain $ch 2>&1 | grep -E "RESULT|FAIL|error" >> gpurun_out/r2a_timing.log
    done
  done
  timeout 120 $B --nc 16 --chain 4 --no-split 2>&1 | grep -E "RESULT|FAIL|error" >> gpurun_out/r2a_timing.log
  timeout 120 $B --nc 16 --chain 4 --no-pdl 2>&1 | grep -E "RESULT|FAIL|error" >> gpurun_out/r2a_timing.log
  timeout 120 $B --nc 16 --chain 4 --slots 32 2>&1 | grep -E "RESULT|FAIL|error" >> gpurun_out/r2a_timing.log
  timeout 120 $B --nc 16 --chain 2 2>&1 | grep -E "RESULT|FAIL|error" >> gpurun_out/r2a_timing.log
  timeout 120 $B --nc 16 --chain 4 --trace > gpurun_out/r2a_trace.log 2>&1
  cat gpurun_out/r2a_timing.log
fi
timeout 900 python -m pytest tests/test_live_reference.py tests/test_ref_names_gpu.py -q -m gpu > gpurun_out/r2a_pytest_new.log 2>&1
tail -15 gpurun_out/r2a_pytest_new.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_all.log 2>&1
tail -5 gpurun_out/r2a_pytest_all.log
