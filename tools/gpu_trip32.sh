#!/usr/bin/env bash
# trip 32: launch list of one train step in the current state (shares), full GPU suite, mrcnn arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v Warning | tail -8 > gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
python tools/ncu_summary.py launches gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -45 gpurun_out/launch_summary.txt
timeout 600 python bench.py --model mrcnn --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mrcnn.json 2> gpurun_out/bench_mrcnn.err; cut -c1-400 gpurun_out/bench_mrcnn.json
timeout 300 python tools/conv_profile.py > gpurun_out/conv_profile.txt 2>&1; head -60 gpurun_out/conv_profile.txt
