#!/bin/bash
# round 2, call F: setmaxnreg (16 consumer warps without spills), split policy, staging in one round trip
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
T=gpurun_out/r2f_timing.log
: > $T
for cfg in "16 2" "8 1" "16 4"; do
  set -- $cfg
  timeout 120 $B --check-only --nc $1 --wps $2 > gpurun_out/r2f_check_$1_$2.log 2>&1; rc=$?
  echo "check nc $1 wps $2 rc=$rc fails=$(grep -c FAIL gpurun_out/r2f_check_$1_$2.log)" | tee -a $T
  if [ $rc = 0 ] && ! grep -q FAIL gpurun_out/r2f_check_$1_$2.log; then
    timeout 120 $B --nc $1 --wps $2 --chain 4 2>&1 | grep -E "RESULT|FAIL|error" >> $T
    timeout 120 $B --nc $1 --wps $2 --chain 1 2>&1 | grep -E "RESULT|FAIL|error" >> $T
  fi
done
grep "stage" gpurun_out/r2f_check_16_2.log | head -8
echo "nocompute:" >> $T
timeout 120 ${B}_nocompute --nc 16 --wps 2 --chain 4 2>&1 | grep -E "RESULT|error" >> $T
timeout 120 $B --nc 16 --wps 2 --chain 4 --trace > gpurun_out/r2f_trace_16_2.log 2>&1
timeout 120 $B --nc 8 --wps 1 --chain 4 --trace > gpurun_out/r2f_trace_8_1.log 2>&1
cat $T
grep -A36 "launch 16" gpurun_out/r2f_trace_16_2.log | head -38
