#!/usr/bin/env bash
# One GPU trip: hardware probe + the GPU parity suite.  Usage (from the repo root, under gpurun): bash tools/gpu_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 120 tools/bin/tc_probe > gpurun_out/tc_probe.txt 2>&1; echo "probe exit $?" >> gpurun_out/tc_probe.txt
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/tc_probe.txt; cat gpurun_out/pytest_gpu.txt
