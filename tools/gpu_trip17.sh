#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q --tb=short -x 2>&1 | tail -5 > gpurun_out/conv_gpu.txt
tail -3 gpurun_out/conv_gpu.txt
MDT_TCW=2 timeout 300 python tools/tcw_prof.py p0_36 c0_18 c1_k7 head64 bb54 > gpurun_out/tcw_prof.txt 2>&1
MDT_TCW=2 PASS=1 timeout 300 python tools/tcw_prof.py p0_36 c1_k7 head64 >> gpurun_out/tcw_prof.txt 2>&1
cat gpurun_out/tcw_prof.txt
