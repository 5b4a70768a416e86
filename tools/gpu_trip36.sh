#!/usr/bin/env bash
# trip 36: train_forward reordered (matching first, detections + D2H last, one sync)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v Warning | tail -6 > gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.json","gpurun_out/bench2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'dom', r.get('ms'), r.get('frac'), 'all', r.get('all_conv_launches',{}).get('conv_ms_per_step'))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 300 python tools/torch_profile.py 2>&1 | tail -4
