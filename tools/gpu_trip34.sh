#!/usr/bin/env bash
# trip 34: wgrad dump without the channel padding, NMS exact-threshold test, full suite, layer bench, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v Warning | tail -6 > gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
PASSES=2 timeout 120 python tools/conv_layer_bench.py p0_36 c0_18 c1_k7 head64 > gpurun_out/wgrad_layers.txt 2>&1; cat gpurun_out/wgrad_layers.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'dom', r.get('ms'), r.get('frac'), 'all', r.get('all_conv_launches',{}).get('conv_ms_per_step'))
    except Exception as e:
        print(f, "ERR", e)
PY
