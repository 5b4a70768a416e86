#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 120 tools/bin/mma_major_probe > gpurun_out/mma_major_probe.txt 2>&1; cat gpurun_out/mma_major_probe.txt
timeout 600 python tools/pw_bench.py 2>&1 | grep "algo 4" > gpurun_out/pw_bench.txt; cat gpurun_out/pw_bench.txt
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
