#!/usr/bin/env bash
# trip 31: new loss/resample kernels (parity), operand-placement probe, wgrad placement experiment, full GPU suite, bench
mkdir -p gpurun_out
timeout 120 tools/bin/mma_major_probe > gpurun_out/mma_major_probe.txt 2>&1; cat gpurun_out/mma_major_probe.txt
timeout 600 python -m pytest tests/test_loss_ops_gpu.py tests/test_consolidate_gpu.py -m gpu -q --tb=short 2>&1 | grep -v Warning | tail -40 > gpurun_out/pytest_loss_ops.txt; cat gpurun_out/pytest_loss_ops.txt
for lay in 0 1; do for gap in 0 16; do
  echo "== MDT_WG_LAYOUT=$lay MDT_WG_GAP=$gap"; MDT_WG_LAYOUT=$lay MDT_WG_GAP=$gap PASSES=2 timeout 120 python tools/conv_layer_bench.py p0_36 c0_18 c1_k7 head64
done; done > gpurun_out/wgrad_layout.txt 2>&1; cat gpurun_out/wgrad_layout.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v Warning | tail -30 > gpurun_out/pytest_gpu.txt; tail -30 gpurun_out/pytest_gpu.txt
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
