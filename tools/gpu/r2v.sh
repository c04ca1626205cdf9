#!/bin/bash
# round 2, call V: the reference README's HF generate() protocol (BASELINE config 3)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=$SECONDS
timeout 1200 python tests/perf/hf_generate_tps.py --which native,ours,ours-graph --iters 5 > gpurun_out/r2v_hf_generate.log 2>&1
echo "took $((SECONDS-S)) s"
tail -6 gpurun_out/r2v_hf_generate.log
