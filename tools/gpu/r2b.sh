#!/bin/bash
# round 2, call B: ring kernel v2 (8-KB slots, warp pairs, lockstep producer lanes)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
T=gpurun_out/r2b_timing.log
: > $T
timeout 120 $B --check-only --nc 16 > gpurun_out/r2b_check.log 2>&1; echo "check rc=$?" >> gpurun_out/r2b_check.log
tail -12 gpurun_out/r2b_check.log
if grep -q FAIL gpurun_out/r2b_check.log || ! grep -q "check rc=0" gpurun_out/r2b_check.log; then
  echo "ring check failed: skipping timing"
else
  for nc in 16 24 8; do
    timeout 120 $B --nc $nc --chain 4 2>&1 | grep -E "RESULT|FAIL|error" >> $T
  done
  timeout 120 $B --nc 16 --chain 1 2>&1 | grep -E "RESULT|FAIL|error" >> $T
  timeout 120 $B --nc 16 --chain 4 --no-split 2>&1 | grep -E "RESULT|FAIL|error" >> $T
  timeout 120 $B --nc 16 --chain 4 --slots 8 2>&1 | grep -E "RESULT|FAIL|error" >> $T
  echo "nocompute:" >> $T
  for nc in 16 8; do
    timeout 120 ${B}_nocompute --nc $nc --chain 4 2>&1 | grep -E "RESULT|error" >> $T
  done
  timeout 120 $B --nc 16 --chain 4 --trace > gpurun_out/r2b_trace.log 2>&1
  cat $T
  grep -A12 "stage 2" gpurun_out/r2b_trace.log | head -30
fi
timeout 600 python -m pytest tests/test_live_reference.py -q -m gpu > gpurun_out/r2b_pytest_new.log 2>&1
tail -5 gpurun_out/r2b_pytest_new.log
