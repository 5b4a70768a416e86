#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -x 2>&1 | grep -v Warning | tail -5
for gnr in 1 0; do
  if [ $gnr = 1 ]; then export MDT_STEM_GENERIC=1; else unset MDT_STEM_GENERIC; fi
  echo "== MDT_STEM_GENERIC=$gnr"; PASSES=02 timeout 100 python tools/conv_layer_bench.py stem
done
unset MDT_STEM_GENERIC
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
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
