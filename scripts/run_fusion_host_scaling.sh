#!/bin/bash
cd "$(dirname "$0")/.."
df -h /tmp /dev/shm | tail -2; nproc; free -g | head -2
for D in /tmp /dev/shm; do
  mkdir -p $D/milb_f
  echo "=== scratch on $D"
  timeout 600 python bench_fusion.py --points 32 --iters 10 --gpus 1,4,8 --same-gpu --modes resident --dir $D/milb_f > gpurun_out/fusion10_$(basename $D).json 2> gpurun_out/fusion10_$(basename $D).err; tail -2 gpurun_out/fusion10_$(basename $D).err
  python - <<PY
import json
for l in open('gpurun_out/fusion10_$(basename $D).json').read().strip().splitlines():
    d = json.loads(l); print(d['n_gpus'], round(d['value'], 3), round(d['steady_state_vols_per_s'], 2), d['resident']['last_time_point_stages'][-2:])
PY
  rm -rf $D/milb_f
done
