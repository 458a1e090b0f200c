#!/bin/bash
# 8-GPU box: spimFusionBatch on 1/2/4/8 GPUs with the harness holding the GPUs' driver state (no persistence daemon on these boxes)
cd "$(dirname "$0")/.."
mkdir -p /dev/shm/milb_f
echo "--- fusion 64 points, regMode 1, 1/2/4/8 GPUs"
timeout 600 python bench_fusion.py --points 64 --iters 10 --gpus 1,2,4,8 --modes resident --dir /dev/shm/milb_f > gpurun_out/fusion15_m1.json 2> gpurun_out/fusion15_m1.err; tail -3 gpurun_out/fusion15_m1.err
echo "--- fusion 32 points, regMode 3 (registration of every time point), 1/8 GPUs"
timeout 600 python bench_fusion.py --points 32 --iters 10 --gpus 1,8 --reg-mode 3 --modes resident --dir /dev/shm/milb_f > gpurun_out/fusion15_m3.json 2> gpurun_out/fusion15_m3.err; tail -3 gpurun_out/fusion15_m3.err
rm -rf /dev/shm/milb_f
python - <<'PY'
import json
for f in ('gpurun_out/fusion15_m1.json', 'gpurun_out/fusion15_m3.json'):
    try:
        for l in open(f).read().strip().splitlines():
            d = json.loads(l); print(f, d['n_gpus'], round(d['value'], 3), round(d['steady_state_vols_per_s'], 2), d['resident']['last_time_point_stages'][-1], d.get('scaling_of_resident'))
    except Exception as e:
        print(f, "parse error", e)
PY
