#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'dom', r.get('ms'), r.get('frac'), 'all', r.get('all_conv_launches',{}).get('conv_ms_per_step'))
        print(d.get('cuda_graph'))
    except Exception as e:
        print(f, "ERR", e)
PY
