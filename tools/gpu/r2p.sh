#!/bin/bash
# round 2, call P: prefill GEMM (packed-weight ring decoupled from the A / activation stages, flexible token tile, one-wave split-K) -- parity, timings
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm" > gpurun_out/r2p_pytest.log 2>&1
tail -3 gpurun_out/r2p_pytest.log
timeout 300 python tools/gemm_check.py 300x1024x4096 1000x1024x4096 1024x14336x4096 700x4096x4096 2048x4096x14336 129x384x512 2>&1 | grep -c OK
GEMM_TIME=1 timeout 900 python tools/gemm_check.py > gpurun_out/r2p_gemm.log 2>&1
grep -c OK gpurun_out/r2p_gemm.log; grep FAIL gpurun_out/r2p_gemm.log | head
grep "fused" gpurun_out/r2p_gemm.log
