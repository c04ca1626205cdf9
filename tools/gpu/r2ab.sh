#!/bin/bash
# round 2, call AB: ncu evidence refreshed for the final code (4 ring launches of 32 stages per step): launch list of the bench command,
# full capture of one ring launch, full capture of the small-batch kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-blockwise --no-sweep --no-tp70b --no-decode --no-graph"
$CMD > gpurun_out/r2ab_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/r2ab_launches.csv $CMD > gpurun_out/r2ab_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemv_ring -s 5 -c 1 -f -o gpurun_out/r2ab_ring $CMD > gpurun_out/r2ab_ncu2.log 2>&1
echo "ring capture rc=$?"
python tools/gpu/tokens_probe.py > gpurun_out/r2ab_tokens_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemv_tokens -s 60 -c 1 -f -o gpurun_out/r2ab_tokens python tools/gpu/tokens_probe.py > gpurun_out/r2ab_ncu3.log 2>&1
echo "tokens capture rc=$?"
tail -2 gpurun_out/r2ab_ncu2.log gpurun_out/r2ab_ncu3.log
