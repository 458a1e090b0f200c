#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/pipe_probe.py 2>&1 | tail -30 | tee gpurun_out/pipe_probe.jsonl
