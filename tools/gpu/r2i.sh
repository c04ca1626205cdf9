#!/bin/bash
# round 2, call I: restructured ring kernel (plans, one-hop exchange, pair-mode SwiGLU) -- correctness, timing, trace
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
T=gpurun_out/r2i_timing.log
: > $T
for cfg in "16 2" "8 1"; do
  set -- $cfg
  timeout 120 $B --check-only --nc $1 --wps $2 > gpurun_out/r2i_check_$1_$2.log 2>&1; rc=$?
  echo "check nc $1 wps $2 rc=$rc fails=$(grep -c FAIL gpurun_out/r2i_check_$1_$2.log)" | tee -a $T
  grep "stage" gpurun_out/r2i_check_$1_$2.log | head -8
  if [ $rc = 0 ] && ! grep -q FAIL gpurun_out/r2i_check_$1_$2.log; then
    timeout 120 $B --nc $1 --wps $2 --chain 4 2>&1 | grep -E "RESULT|FAIL|error" >> $T
    timeout 120 $B --nc $1 --wps $2 --chain 1 2>&1 | grep -E "RESULT|FAIL|error" >> $T
  fi
done
timeout 120 $B --check-only --nc 16 --wps 2 --no-split > gpurun_out/r2i_check_ns.log 2>&1; echo "no-split check rc=$? fails=$(grep -c FAIL gpurun_out/r2i_check_ns.log)" | tee -a $T
echo "nocompute:" >> $T
timeout 120 ${B}_nocompute --nc 16 --wps 2 --chain 4 2>&1 | grep -E "RESULT|error" >> $T
timeout 120 $B --nc 16 --wps 2 --chain 4 --trace > gpurun_out/r2i_trace_16_2.log 2>&1
cat $T
grep -A40 "launch 16" gpurun_out/r2i_trace_16_2.log | head -44
