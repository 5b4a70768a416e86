#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_model_golden.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 280 python tools/cudnn_convs_bench.py --steps 5 --warmup 2 --kinds cudnn_fp32 cudnn_tf32 > gpurun_out/cudnn_convs.jsonl 2> gpurun_out/cudnn_convs.err; cat gpurun_out/cudnn_convs.jsonl; tail -3 gpurun_out/cudnn_convs.err
