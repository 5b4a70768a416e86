#!/usr/bin/env bash
# One GPU trip.  Usage (repo root, under gpurun): bash tools/gpu_check.sh [bench args]   env: SKIP_MICRO=1, SKIP_TESTS=1, PROFILE=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 --durations=12 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt | cut -c1-300
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.txt
fi
if [ -z "$SKIP_MICRO" ]; then timeout 600 python tools/microbench.py > gpurun_out/microbench.json 2> gpurun_out/microbench.err; tail -3 gpurun_out/microbench.err; fi
if [ -n "$PROFILE" ]; then timeout 600 python tools/conv_profile.py > gpurun_out/conv_profile.txt 2>&1; cat gpurun_out/conv_profile.txt | head -60; fi
timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-1200; tail -5 gpurun_out/bench.err
