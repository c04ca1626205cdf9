#!/bin/bash
# round 2, call C: ring kernel v3 (fixed slot ownership, all CTAs active, private x) + ncu capture
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
T=gpurun_out/r2c_timing.log
: > $T
ok=1
for nc in 16 8; do
  timeout 120 $B --check-only --nc $nc > gpurun_out/r2c_check_$nc.log 2>&1; echo "check rc=$?" >> gpurun_out/r2c_check_$nc.log
  tail -8 gpurun_out/r2c_check_$nc.log
  if grep -q FAIL gpurun_out/r2c_check_$nc.log || ! grep -q "check rc=0" gpurun_out/r2c_check_$nc.log; then ok=0; fi
done
timeout 120 $B --check-only --nc 16 --slots 8 > gpurun_out/r2c_check_s8.log 2>&1; echo "check rc=$?" >> gpurun_out/r2c_check_s8.log; tail -3 gpurun_out/r2c_check_s8.log
timeout 120 $B --check-only --nc 16 --no-split > gpurun_out/r2c_check_ns.log 2>&1; echo "check rc=$?" >> gpurun_out/r2c_check_ns.log; tail -3 gpurun_out/r2c_check_ns.log
if [ $ok = 0 ]; then
  echo "ring check failed: skipping timing"
else
  for args in "--nc 16 --chain 4" "--nc 16 --chain 1" "--nc 16 --chain 4 --slots 8" "--nc 16 --chain 4 --no-split" "--nc 8 --chain 4" "--nc 24 --chain 4"; do
    timeout 120 $B $args 2>&1 | grep -E "RESULT|FAIL|error" >> $T
  done
  echo "nocompute:" >> $T
  for args in "--nc 16 --chain 4" "--nc 8 --chain 4"; do
    timeout 120 ${B}_nocompute $args 2>&1 | grep -E "RESULT|error" >> $T
  done
  timeout 120 $B --nc 16 --chain 4 --trace > gpurun_out/r2c_trace.log 2>&1
  timeout 120 ${B}_nocompute --nc 16 --chain 4 --trace > gpurun_out/r2c_trace_nocompute.log 2>&1
  cat $T
  timeout 300 $B --nc 16 --chain 4 --layers 2 --pool 2 --iters 1 > gpurun_out/r2c_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_ring -s 2 -c 1 -f -o gpurun_out/r2c_ring \
     $B --nc 16 --chain 4 --layers 2 --pool 2 --iters 1 > gpurun_out/r2c_ncu.log 2>&1
  tail -3 gpurun_out/r2c_ncu.log
fi
