#!/bin/bash
# 4-GPU box: the driver's scaling command for N = 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/bench_p4.json 2> gpurun_out/bench_p4.err
tail -2 gpurun_out/bench_p4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_p4.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'] and d['e2e']['value'])
c = d.get('config3'); print('config3', c.get('ms_per_iteration'), c.get('bit_identical_to_single_gpu'), c.get('roofline', {}).get('frac'), c.get('error'))
PY
