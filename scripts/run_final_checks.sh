#!/bin/bash
# What the driver runs at round end, on one GPU: the GPU test suite, smoke(), both bench arms.
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; tail -c 600 gpurun_out/bench_final_reference.json; echo
python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
r = d['roofline']
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')})
print('roofline', {k: r[k] for k in ('achieved', 'peak', 'frac', 'traffic', 'traffic_per_voxel', 'frac_of_nominal_8TBps')})
print('kernels', [(k['kernel'][:12], round(k['ms_per_launch'], 4), round(k['frac_of_peak'], 3)) for k in r['kernels']])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('parity', d['parity'])
print('cpu', d['cpu_baseline']['value'], 'refgpu', d['reference_gpu_yardstick'])
print('yard', d['yardstick'])
print('reg', d['registration'])
print('config2', d['config2']['value'], d['config2']['roofline']['frac'])
print('config3', d['config3'])
PY
