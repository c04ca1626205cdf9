#!/bin/bash
# round 2, call AA: ncu evidence -- launch list of the bench command, full capture of the ring GEMV, full capture of the prefill GEMM
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-blockwise --no-sweep --no-tp70b --no-decode --no-graph"
$CMD > gpurun_out/r2aa_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 40 -c 120 --csv --log-file gpurun_out/r2aa_launches.csv $CMD > gpurun_out/r2aa_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemv_ring -s 6 -c 1 -f -o gpurun_out/r2aa_ring $CMD > gpurun_out/r2aa_ncu2.log 2>&1
echo "ring capture rc=$?"
python tools/gpu/gemm_probe.py > gpurun_out/r2aa_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_dequant -s 7 -c 1 -f -o gpurun_out/r2aa_gemm python tools/gpu/gemm_probe.py > gpurun_out/r2aa_ncu3.log 2>&1
echo "gemm capture rc=$?"
cat gpurun_out/r2aa_gemm_plain.log
tail -2 gpurun_out/r2aa_ncu2.log
