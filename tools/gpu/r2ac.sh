#!/bin/bash
# round 2, call AC: ncu launch list of the bench command (GEMV kernels only), final code
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-blockwise --no-sweep --no-tp70b --no-decode --no-graph"
$CMD > gpurun_out/r2ac_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:gemv -c 170 --csv --log-file gpurun_out/r2ac_launches.csv $CMD > gpurun_out/r2ac_ncu1.log 2>&1
echo "launch list rc=$?"
