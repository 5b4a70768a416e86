#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_golden.py tests/test_conv_gpu.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt
timeout 300 python tools/conv_profile.py > gpurun_out/conv_profile.txt 2>&1
head -60 gpurun_out/conv_profile.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -32 gpurun_out/launch_summary.txt
