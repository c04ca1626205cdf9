#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="300x1024x4096 256x1024x4096 512x1024x4096 257x1024x4096 256x128x4096 256x128x512 64x1024x14336 1024x4096x4096"
echo "== default"; timeout 300 python tools/gemm_check.py $C 2>&1 | grep "float16 " | grep "fp4"
echo "== splits 1"; Q4_GEMM_SPLITS=1 timeout 300 python tools/gemm_check.py $C 2>&1 | grep "float16 " | grep "fp4"
echo "== splits 2"; Q4_GEMM_SPLITS=2 timeout 300 python tools/gemm_check.py $C 2>&1 | grep "float16 " | grep "fp4"
