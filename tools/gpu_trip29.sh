#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 120 tools/bin/mma_pipe_probe 2>&1 | grep " 3600 \| 36      0      0      0      0" > gpurun_out/mma_pipe_probe_again.txt; cat gpurun_out/mma_pipe_probe_again.txt
timeout 600 python tools/pw_bench.py 2>&1 | grep "algo 4" > gpurun_out/pw_bench.txt; cat gpurun_out/pw_bench.txt
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
