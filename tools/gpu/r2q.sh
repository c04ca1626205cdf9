#!/bin/bash
# round 2, call Q: full GPU suite, prefill GEMM table, bench -- after the GEMM pipeline change and the small-batch kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest_all.log 2>&1
tail -4 gpurun_out/r2q_pytest_all.log
GEMM_TIME=1 timeout 900 python tools/gemm_check.py > gpurun_out/r2q_gemm.log 2>&1
grep -c OK gpurun_out/r2q_gemm.log; grep FAIL gpurun_out/r2q_gemm.log | head -3
grep fused gpurun_out/r2q_gemm.log | awk '{print $1,$2,$3,$4,$5,$7,$12}' | head -30
