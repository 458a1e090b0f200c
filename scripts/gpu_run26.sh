#!/bin/bash
cd "$(dirname "$0")/.."
for cap in 0 111 74 56 40 20; do MILB_GRID_CAP=$cap timeout 120 python scripts/cap_probe.py 2>&1 | tail -1; done | tee gpurun_out/cap_probe.jsonl
