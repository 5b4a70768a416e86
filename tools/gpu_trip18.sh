#!/usr/bin/env bash
mkdir -p gpurun_out
for sk in 0 1 2; do
  echo "== MDT_WG_SKIP=$sk" >> gpurun_out/wg_skip.txt
  MDT_WG_SKIP=$sk PASSES=2 REPS=5 timeout 300 python tools/conv_layer_bench.py p0_36 c0_18 c1_k7 head64 >> gpurun_out/wg_skip.txt 2>&1
done
cat gpurun_out/wg_skip.txt
