#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_loss_ops_gpu.py tests/test_model_golden.py -m gpu -q --tb=short -x 2>&1 | grep -v Warning | tail -3
timeout 200 python tools/loss_bench.py > gpurun_out/loss_bench.json 2> gpurun_out/loss_bench.err; cat gpurun_out/loss_bench.json; tail -3 gpurun_out/loss_bench.err
