#!/bin/bash
# sustained-mode A/B of two builds: bench.py (3 + 6 steps of 50 iterations at 512^3), default library vs lib/libapi_<tag>.so
cd "$(dirname "$0")/.."
TAG=${1:-f32x2}
for rep in 1 2; do
for lib in default $TAG; do
  if [ $lib = default ]; then unset MILB_LIBAPI; else export MILB_LIBAPI=$PWD/microimagelib_b200/lib/libapi_$lib.so; fi
  python bench.py --gpus 1 --steps 6 --warmup 3 --no-traffic --no-refgpu --no-config3 --no-cpu-baseline --no-yardstick --no-e2e --shape2 "" > gpurun_out/ab_$lib.json 2> gpurun_out/ab_$lib.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/ab_$lib.json').read().strip().splitlines()[-1])
print('$lib', round(d['ms_per_step'], 2), 'ms/step frac', round(d['roofline']['frac'], 4), 'alone', round(d['roofline']['timed_alone']['ms_per_launch'], 4), d['clocks'])
PY
done
done
