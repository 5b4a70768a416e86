#!/usr/bin/env bash
# One GPU trip: GPU parity suite + smoke + microbench + a short bench.  Usage (repo root, under gpurun): bash tools/gpu_check.sh [bench args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.txt
timeout 600 python tools/microbench.py > gpurun_out/microbench.json 2> gpurun_out/microbench.err; tail -3 gpurun_out/microbench.err
timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
