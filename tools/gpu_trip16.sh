#!/usr/bin/env bash
mkdir -p gpurun_out
for sk in 0 1 2; do
  echo "== MDT_TCW_SKIP=$sk" >> gpurun_out/tcw_skip.txt
  MDT_TCW=2 MDT_TCW_SKIP=$sk timeout 300 python tools/tcw_prof.py p0_36 c0_18 head64 >> gpurun_out/tcw_skip.txt 2>&1
done
echo "== SKIP=2 NBUF... TL=1" >> gpurun_out/tcw_skip.txt
MDT_TCW=2 MDT_TCW_SKIP=2 MDT_TCW_TL=1 timeout 300 python tools/tcw_prof.py p0_36 >> gpurun_out/tcw_skip.txt 2>&1
cat gpurun_out/tcw_skip.txt
