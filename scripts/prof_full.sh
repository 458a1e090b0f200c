#!/bin/bash
# ncu --set full capture of the RL loop kernels (one launch each), report -> gpurun_out/prof_<tag>.ncu-rep
cd "$(dirname "$0")/.."
TAG=${1:-r1}
export PROBE_ITERS=2 PROBE_CHUNK=${PROBE_CHUNK:-0}
# skip the OTF-generation launches: profile launches from the iteration loop (-s skips first N matching)
ncu --set full --clock-control none --import-source on -k regex:'k_ypassT|k_zconvT|k_ypassF|k_xpassF' -s ${PROBE_SKIP:-60} -c ${PROBE_COUNT:-8} \
    -o gpurun_out/prof_$TAG -f python scripts/prof_run.py > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/prof_full.log
ls -la gpurun_out/*.ncu-rep
