#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
for f in ("gpurun_out/bench_n2.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, 'e2e', d['e2e']['ms_per_step'])
    except Exception as e:
        print(f, "ERR", e)
PY
