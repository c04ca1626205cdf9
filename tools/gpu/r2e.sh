#!/bin/bash
# round 2, call E: what bounds the consumer loop?  (skip-lookup / no-MMA experiments)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
T=gpurun_out/r2e_timing.log
: > $T
for v in "" _skip _nomma _nocompute; do
  echo "variant '$v'" >> $T
  for cfg in "8 1" "16 2"; do
    set -- $cfg
    timeout 120 ${B}$v --nc $1 --wps $2 --chain 4 2>&1 | grep -E "RESULT|error" >> $T
  done
  timeout 120 ${B}$v --nc 8 --wps 1 --chain 4 --trace 2>&1 | grep -A36 "launch 16" | grep -E "stage|x staged|loop end" > gpurun_out/r2e_trace$v.log
done
cat $T
for v in "" _skip _nomma; do echo "== $v"; cat gpurun_out/r2e_trace$v.log | head -14; done
