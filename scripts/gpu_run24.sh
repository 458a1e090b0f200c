#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/zrow_probe.py 2>&1 | grep -v rel_l2 | tee gpurun_out/zrow_probe_timing.jsonl
timeout 900 python -m pytest tests/test_gpu_decon.py tests/test_gpu_reference_pinned.py tests/test_reference_golden.py -m gpu -x -q 2>&1 | tail -4
