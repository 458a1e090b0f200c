#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/zrow_probe.py 2>&1 | grep -v rel_l2 | tee gpurun_out/zrow_probe_timing.jsonl
