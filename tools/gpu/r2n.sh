#!/bin/bash
# round 2, call N: ncu --set full of the prefill GEMM after the pipeline change (M = 4096, 14336x4096)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python tools/gpu/gemm_probe.py > gpurun_out/r2n_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_dequant -s 7 -c 1 -f -o gpurun_out/r2n_gemm python tools/gpu/gemm_probe.py > gpurun_out/r2n_ncu.log 2>&1
echo "gemm capture rc=$?"
cat gpurun_out/r2n_gemm_plain.log
tail -2 gpurun_out/r2n_ncu.log
ls -la gpurun_out/r2n_gemm.ncu-rep
