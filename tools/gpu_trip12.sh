#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_golden.py -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
PASSES=012 REPS=5 timeout 300 python tools/conv_layer_bench.py p0_36 c0_18 c1_k7 head64 head64_p3 > gpurun_out/layer_bench.txt 2>&1
cat gpurun_out/layer_bench.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'])"
