#!/usr/bin/env bash
mkdir -p gpurun_out
for ch in 512 1024 2048 4096; do
  echo "== MDT_WG_CHAIN=$ch" >> gpurun_out/wg_chain.txt
  MDT_WG_CHAIN=$ch timeout 300 python tools/wgrad_precision.py 2>&1 | grep "128, 128, 128\|64, 64, 128" >> gpurun_out/wg_chain.txt
  MDT_WG_CHAIN=$ch PASSES=2 REPS=5 timeout 300 python tools/conv_layer_bench.py p0_36 c0_18 c1_k7 head64 >> gpurun_out/wg_chain.txt 2>&1
done
cat gpurun_out/wg_chain.txt
