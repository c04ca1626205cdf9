#!/bin/bash
# round 2, call W: host cost per Linear4bit call, then the module copy / pickle test
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/gpu/host_cost.py > gpurun_out/r2w_host_cost.log 2>&1
head -40 gpurun_out/r2w_host_cost.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "copies_and_pickles or linear4bit or group" > gpurun_out/r2w_pytest.log 2>&1
tail -4 gpurun_out/r2w_pytest.log
