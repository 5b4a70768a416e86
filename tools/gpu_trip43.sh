#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -x -k "full_resolution" --durations=3 2>&1 | grep -v Warning | tail -8
timeout 300 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_conv_gpu.py::test_full_resolution_layers_vs_fp64 2>&1 | grep -v Warning | tail -3
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'traffic', d['roofline']['traffic'])
PY
