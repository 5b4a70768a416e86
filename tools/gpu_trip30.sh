#!/usr/bin/env bash
mkdir -p gpurun_out
for sk in 0 1 2 3 4; do echo "== MDT_TCW_SKIP=$sk"; MDT_TCW_SKIP=$sk timeout 120 python tools/tcw_prof.py p0_36 c1_k7 head64; done > gpurun_out/tcw_skip2.txt 2>&1
cat gpurun_out/tcw_skip2.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v Warning | tail -30 > gpurun_out/pytest_gpu.txt
tail -30 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --model mrcnn --steps 10 --warmup 3 > gpurun_out/bench_mrcnn.json 2> gpurun_out/bench_mrcnn.err
python - <<'PY'
import json
for f in ("gpurun_out/bench.json", "gpurun_out/bench_mrcnn.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'dom', r.get('ms'), r.get('frac'), 'all', r.get('all_conv_launches',{}).get('conv_ms_per_step'))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 300 python tools/conv_profile.py > gpurun_out/conv_profile.txt 2>&1
head -30 gpurun_out/conv_profile.txt
