#!/bin/bash
# final refresh of the round's evidence on one GPU: profiles (ncu), then what the driver runs (tests, smoke, both bench arms)
cd "$(dirname "$0")/.."
bash scripts/make_profiles.sh r02 > gpurun_out/make_profiles.log 2>&1; tail -2 gpurun_out/make_profiles.log
bash scripts/run_final_checks.sh
