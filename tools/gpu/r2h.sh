#!/bin/bash
# round 2, call H: ncu capture of the ring kernel (stall reasons, pipe utilisation) + experiment variants
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tools/micro/_bin/ring_bench
T=gpurun_out/r2h_timing.log
: > $T
for v in _skip _nomma; do
  echo "variant '$v'" >> $T
  for cfg in "8 1" "16 2"; do
    set -- $cfg
    [ -x ${B}$v ] && timeout 120 ${B}$v --nc $1 --wps $2 --chain 4 2>&1 | grep -E "RESULT|error" >> $T
  done
done
cat $T
timeout 300 $B --nc 16 --wps 2 --chain 4 --layers 2 --pool 2 --iters 1 > gpurun_out/r2h_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_ring -s 2 -c 1 -f -o gpurun_out/r2h_ring \
   $B --nc 16 --wps 2 --chain 4 --layers 2 --pool 2 --iters 1 > gpurun_out/r2h_ncu.log 2>&1
tail -3 gpurun_out/r2h_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_ring -s 2 -c 1 -f -o gpurun_out/r2h_ring81 \
   $B --nc 8 --wps 1 --chain 4 --layers 2 --pool 2 --iters 1 > gpurun_out/r2h_ncu81.log 2>&1
tail -3 gpurun_out/r2h_ncu81.log
