#!/bin/bash
# 8-GPU box: the driver's scaling command for N = 8 (weak scaling + config 3 on 8 GPUs) and spimFusionBatch on 1/2/4/8 GPUs
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench9_p8.json 2> gpurun_out/bench9_p8.err
tail -2 gpurun_out/bench9_p8.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench9_p8.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'] and d['e2e']['value'])
    print('config3', json.dumps(d.get('config3'))[:1200])
except Exception as e:
    print("bench parse error", e)
PY
echo "--- fusion 64 points, regMode 1, 1/2/4/8 GPUs"
timeout 900 python bench_fusion.py --points 64 --iters 10 --gpus 1,2,4,8 --modes resident > gpurun_out/fusion9_m1.json 2> gpurun_out/fusion9_m1.err; tail -3 gpurun_out/fusion9_m1.err
echo "--- fusion 32 points, regMode 3 (registration of every time point), 1/8 GPUs"
timeout 900 python bench_fusion.py --points 32 --iters 10 --gpus 1,8 --reg-mode 3 --modes resident > gpurun_out/fusion9_m3.json 2> gpurun_out/fusion9_m3.err; tail -3 gpurun_out/fusion9_m3.err
python - <<'PY'
import json
for f in ('gpurun_out/fusion9_m1.json', 'gpurun_out/fusion9_m3.json'):
    try:
        for l in open(f).read().strip().splitlines():
            d = json.loads(l); print(f, d['n_gpus'], round(d['value'], 3), round(d['steady_state_vols_per_s'], 2), d.get('scaling_of_resident'))
    except Exception as e:
        print(f, "parse error", e)
PY
