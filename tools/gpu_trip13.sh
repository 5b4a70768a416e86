#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -6 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --model mrcnn --steps 8 --warmup 3 > gpurun_out/bench_mrcnn.json 2> gpurun_out/bench_mrcnn.err
tail -3 gpurun_out/bench_mrcnn.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.json", "gpurun_out/bench_mrcnn.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('metric','value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['ms'] if d.get('roofline') else None,
              d['roofline'].get('slowest_launch') if d.get('roofline') else None)
    except Exception as e:
        print(f, "ERR", e)
PY
