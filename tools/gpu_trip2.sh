#!/usr/bin/env bash
# GPU trip: conv parity (tap-stacked kernel), layer A/B timings, reference-pinned model tests with full tracebacks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q --tb=short 2>&1 | tail -80 > gpurun_out/conv_gpu.txt
tail -5 gpurun_out/conv_gpu.txt
for v in 1 0; do
  echo "== MDT_TCW=$v" >> gpurun_out/layer_ab.txt
  MDT_TCW=$v PASSES=01 REPS=5 timeout 300 python tools/conv_layer_bench.py p0_36 c0_18 c1_k7 head64 >> gpurun_out/layer_ab.txt 2>&1
done
cat gpurun_out/layer_ab.txt
timeout 600 python -m pytest tests/test_model_golden.py -m gpu -q --tb=short 2>&1 > gpurun_out/golden_gpu.txt
tail -15 gpurun_out/golden_gpu.txt
