#!/usr/bin/env bash
# Evidence trip (under gpurun, repo root): tests, smoke, microbench, ncu launch list of one train step, one `ncu --set full` capture of every
# hot-path kernel at its BASELINE shape (tools/ncu_ops.py), then un-profiled bench lines (both model arms).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 --durations=8 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt; cut -c1-160 gpurun_out/pytest_gpu.txt | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
[ -z "$SKIP_MICRO" ] && timeout 400 python tools/microbench.py > gpurun_out/microbench.json 2> gpurun_out/microbench.err; tail -2 gpurun_out/microbench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
python tools/ncu_summary.py launches gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -14 gpurun_out/launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k 'regex:conv_|nms_|roi_|match_|split_rows|pack_weights|wgrad|bias_grad|stem_' -f -o gpurun_out/ops \
    python tools/ncu_ops.py > gpurun_out/ncu_ops.log 2>&1; tail -3 gpurun_out/ncu_ops.log | cut -c1-200; ls -la gpurun_out/ops.ncu-rep
ncu -i gpurun_out/ops.ncu-rep --page raw --csv > gpurun_out/ops_raw.csv 2>/dev/null; wc -c gpurun_out/ops_raw.csv
python tools/ncu_summary.py ops gpurun_out/ops_raw.csv > gpurun_out/ops_summary.txt 2>&1; head -50 gpurun_out/ops_summary.txt
[ "$(stat -c %s gpurun_out/ops.ncu-rep 2>/dev/null || echo 0)" -gt 30000000 ] && rm -f gpurun_out/ops.ncu-rep   # gpurun_out is capped at 64 MiB
timeout 300 python tools/conv_profile.py > gpurun_out/conv_profile.txt 2>&1; head -12 gpurun_out/conv_profile.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --model mrcnn --steps 20 --warmup 5 > gpurun_out/bench_mrcnn.json 2> gpurun_out/bench_mrcnn.err; cut -c1-300 gpurun_out/bench_mrcnn.json
MDT_REF_BUDGET_S=70 timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 200 python tools/loss_bench.py > gpurun_out/loss_bench.json 2> gpurun_out/loss_bench.err; head -c 600 gpurun_out/loss_bench.json
[ -n "$WITH_CUDNN" ] && timeout 400 python tools/cudnn_convs_bench.py --steps 5 --warmup 2 --kinds cudnn_fp32 cudnn_tf32 > gpurun_out/cudnn_convs.jsonl 2> gpurun_out/cudnn_convs.err; cat gpurun_out/cudnn_convs.jsonl 2>/dev/null
