#!/bin/bash
# 8-GPU box: the driver's scaling command for N = 8 (weak scaling of independent volumes + config 3: one 1024x1024x512 pair on 8 GPUs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_p8.json 2> gpurun_out/bench_p8.err
tail -2 gpurun_out/bench_p8.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_p8.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'] and d['e2e']['value'])
    print('config3', json.dumps(d.get('config3'))[:1200])
except Exception as e:
    print("bench parse error", e)
PY
