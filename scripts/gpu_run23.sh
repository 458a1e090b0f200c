#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "--- gpu tests"
python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "--- bench"
python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench23.json 2> gpurun_out/bench23.err; tail -3 gpurun_out/bench23.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench23.json').read().strip().splitlines()[-1])
r = d['roofline']
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')})
print('roofline', {k: r[k] for k in ('achieved', 'peak', 'frac', 'traffic', 'traffic_per_voxel', 'frac_of_nominal_8TBps')})
print('kernels', [(k['kernel'][:12], round(k['ms_per_launch'], 4), round(k['frac_of_peak'], 3)) for k in r['kernels']])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('parity', d['parity'])
print('config2', d['config2']['value'], d['config2']['roofline']['frac'])
print('config3', d['config3'])
PY
echo "--- profiles"
bash scripts/make_profiles.sh r02 > gpurun_out/make_profiles.log 2>&1; tail -3 gpurun_out/make_profiles.log
