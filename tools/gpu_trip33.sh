#!/usr/bin/env bash
# trip 33: NMS division filter (opt-in parity + A/B), bias-gradient fix, bench
mkdir -p gpurun_out; rm -f gpurun_out/nms_filter_ab.txt
bash tools/gpu_experiments.sh 2>&1 | tail -20
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -k "bias_gradient or fused_backward" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
MDT_NMS_FILTER=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_filter.json 2> gpurun_out/bench_filter.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.json","gpurun_out/bench_filter.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'dom', r.get('ms'), r.get('frac'), 'all', r.get('all_conv_launches',{}).get('conv_ms_per_step'))
    except Exception as e:
        print(f, "ERR", e)
PY
