#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v Warning | tail -5 > gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --graphs > gpurun_out/bench_graphs.json 2> gpurun_out/bench_graphs.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.json","gpurun_out/bench_graphs.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'dom', r.get('ms'), r.get('frac'), 'all', r.get('all_conv_launches',{}).get('conv_ms_per_step'))
        print(d.get('cuda_graph'))
    except Exception as e:
        print(f, "ERR", e)
PY
