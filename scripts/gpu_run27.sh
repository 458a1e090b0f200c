#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python scripts/pipe_probe.py "0.27,0.46" "0.29,0.43" 2>&1 | tail -12 | tee gpurun_out/pipe_probe.jsonl
PROBE_SHAPE=512,512,512 timeout 120 python scripts/cap_probe.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_decon.py -m gpu -x -q 2>&1 | tail -3
