#!/usr/bin/env bash
mkdir -p gpurun_out
PASS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_wgrad_kernel -c 1 -o gpurun_out/ncu_wgrad_p0 python tools/tcw_prof.py p0_36 --once > gpurun_out/ncu_wgrad.log 2>&1
tail -2 gpurun_out/ncu_wgrad.log
PASS=0 MDT_TCW=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 -o gpurun_out/ncu_tc_head64 python tools/tcw_prof.py head64 --once > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
