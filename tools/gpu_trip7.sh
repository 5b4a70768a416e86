#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q --tb=short -x 2>&1 | tail -30 > gpurun_out/conv_gpu.txt
tail -4 gpurun_out/conv_gpu.txt
timeout 300 python tools/tcw_prof.py p0_36 c0_18 c1_k7 head64 bb54 > gpurun_out/tcw_prof.txt 2>&1
PASS=1 timeout 300 python tools/tcw_prof.py p0_36 c1_k7 head64 >> gpurun_out/tcw_prof.txt 2>&1
echo "== STACK=1" >> gpurun_out/tcw_prof.txt
MDT_TCW_STACK=1 timeout 300 python tools/tcw_prof.py p0_36 >> gpurun_out/tcw_prof.txt 2>&1
echo "== STACK=0" >> gpurun_out/tcw_prof.txt
MDT_TCW_STACK=0 timeout 300 python tools/tcw_prof.py c0_18 >> gpurun_out/tcw_prof.txt 2>&1
cat gpurun_out/tcw_prof.txt
