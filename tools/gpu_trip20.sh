#!/usr/bin/env bash
mkdir -p gpurun_out
MDT_TCW=2 MDT_TCW_NARROW=1 timeout 600 python -m pytest tests/test_conv_gpu.py -q --tb=short -x 2>&1 | tail -3
echo "== NARROW=1" > gpurun_out/tcw_prof2.txt
MDT_TCW=2 MDT_TCW_NARROW=1 timeout 300 python tools/tcw_prof.py c0_18 c1_k7 >> gpurun_out/tcw_prof2.txt 2>&1
MDT_TCW=2 MDT_TCW_NARROW=1 PASS=1 timeout 300 python tools/tcw_prof.py c0_18 c1_k7 >> gpurun_out/tcw_prof2.txt 2>&1
echo "== STACK=0 (c0_18), STACK=1 (p0_36)" >> gpurun_out/tcw_prof2.txt
MDT_TCW=2 MDT_TCW_STACK=0 timeout 300 python tools/tcw_prof.py c0_18 >> gpurun_out/tcw_prof2.txt 2>&1
MDT_TCW=2 MDT_TCW_STACK=1 timeout 300 python tools/tcw_prof.py p0_36 >> gpurun_out/tcw_prof2.txt 2>&1
echo "== old kernel (MDT_TCW=0) for comparison, presplit input" >> gpurun_out/tcw_prof2.txt
MDT_TCW=0 timeout 300 python tools/tcw_prof.py p0_36 c0_18 c1_k7 head64 bb54 >> gpurun_out/tcw_prof2.txt 2>&1
MDT_TCW=0 PASS=1 timeout 300 python tools/tcw_prof.py p0_36 c0_18 c1_k7 head64 >> gpurun_out/tcw_prof2.txt 2>&1
MDT_TCW=2 PASS=1 timeout 300 python tools/tcw_prof.py c0_18 bb54 >> gpurun_out/tcw_prof2.txt 2>&1
cat gpurun_out/tcw_prof2.txt | cut -c1-60
