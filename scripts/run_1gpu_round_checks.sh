#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- gpu tests"
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/gpu_tests11.log; tail -12 gpurun_out/gpu_tests11.log
echo "--- non-pow2 boxes"
python scripts/nonpow2_probe.py 2>&1 | tail -8
echo "--- zncc"
python scripts/zncc_probe.py 2>&1 | head -7
echo "--- config 4 registration"
python scripts/config4_reg.py 2>&1 | tail -3
echo "--- profiles"
bash scripts/make_profiles.sh r02 > gpurun_out/make_profiles.log 2>&1; tail -3 gpurun_out/make_profiles.log
