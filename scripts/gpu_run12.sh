#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- gpu tests"
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/gpu_tests12.log; tail -12 gpurun_out/gpu_tests12.log
echo "--- zncc"
python scripts/zncc_probe.py 2>&1 | head -6
echo "--- fusion, 1 GPU, 24 points"
python bench_fusion.py --points 24 --iters 10 --gpus 1 --modes resident --dir /dev/shm/milb_f > gpurun_out/fusion12_p1.json 2> gpurun_out/fusion12_p1.err; tail -2 gpurun_out/fusion12_p1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/fusion12_p1.json').read().strip().splitlines()[-1]); print(d['value'], d['steady_state_vols_per_s'], d['resident']['last_time_point_stages'])
PY
echo "--- fusion dispim + 3-D MIPs"
python bench_fusion.py --points 8 --iters 10 --gpus 1 --dispim --mip3d --modes resident --dir /dev/shm/milb_f > gpurun_out/fusion12_dispim.json 2> gpurun_out/fusion12_dispim.err; tail -2 gpurun_out/fusion12_dispim.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/fusion12_dispim.json').read().strip().splitlines()[-1]); print(d['value'], d['steady_state_vols_per_s'], d['resident']['last_time_point_stages'])
PY
echo "--- smoke"
python __graft_entry__.py smoke 2>&1 | tail -2
