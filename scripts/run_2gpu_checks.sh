#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- 2-GPU distributed parity test"
python -m pytest tests/test_gpu_dist_decon.py -q 2>&1 | tail -5
echo "--- bench.py --gpus 2 (weak scaling + config 3 on 2 GPUs)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench8_p2.json 2> gpurun_out/bench8_p2.err
tail -3 gpurun_out/bench8_p2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench8_p2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'] and d['e2e']['value'])
print('config3', json.dumps(d.get('config3'))[:1500])
PY
echo "--- bench.py --impl reference under torchrun"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
echo "--- fusion 1,2 GPUs"
python bench_fusion.py --points 24 --iters 10 --gpus 1,2 --modes resident > gpurun_out/fusion8.json 2> gpurun_out/fusion8.err; tail -3 gpurun_out/fusion8.err
python - <<'PY'
import json
for l in open('gpurun_out/fusion8.json').read().strip().splitlines():
    d = json.loads(l); print(d['n_gpus'], d['value'], d['steady_state_vols_per_s'], d.get('scaling_of_resident'))
PY
