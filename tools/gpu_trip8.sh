#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tcw_kernel -c 1 -o gpurun_out/ncu_tcw_p0_v5 python tools/tcw_prof.py p0_36 --once > gpurun_out/ncu_tcw.log 2>&1
tail -2 gpurun_out/ncu_tcw.log
