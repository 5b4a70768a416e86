#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/torch_profile.py > gpurun_out/torch_profile.txt 2>&1; tail -42 gpurun_out/torch_profile.txt | cut -c1-200
timeout 600 python bench.py --model mrcnn --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mrcnn.json 2> gpurun_out/bench_mrcnn.err; cut -c1-330 gpurun_out/bench_mrcnn.json
timeout 600 python -m pytest tests/test_model_golden.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -3
