#!/bin/bash
# round 2, call T: the small-batch tokens kernel -- tests, then timings
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tokens_gpu.py -x -q > gpurun_out/r2t_pytest.log 2>&1
tail -5 gpurun_out/r2t_pytest.log
timeout 600 python tools/gpu/tokens_probe.py > gpurun_out/r2t_probe.log 2>&1
tail -8 gpurun_out/r2t_probe.log
