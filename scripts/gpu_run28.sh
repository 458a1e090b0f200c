#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- 2-GPU distributed parity test"
timeout 600 python -m pytest tests/test_gpu_dist_decon.py -q 2>&1 | tail -5
echo "--- bench.py --gpus 2 (weak scaling + config 3 on 2 GPUs)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench28_p2.json 2> gpurun_out/bench28_p2.err
tail -3 gpurun_out/bench28_p2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench28_p2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'] and d['e2e']['value'])
print('config3', json.dumps(d.get('config3'))[:1500])
PY
