#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/pw_bench.py > gpurun_out/pw_bench.txt 2>&1; cat gpurun_out/pw_bench.txt
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
