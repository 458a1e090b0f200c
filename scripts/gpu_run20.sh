#!/bin/bash
# row convolution: agreement with the transposing kernels, timings, then the decon GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "--- zrow probe"
timeout 600 python scripts/zrow_probe.py 2>&1 | tee gpurun_out/zrow_probe.jsonl
echo "--- decon tests"
timeout 900 python -m pytest tests/test_gpu_decon.py tests/test_gpu_reference_pinned.py -m gpu -x -q -k "decon or fast or dual or single" 2>&1 | tail -8
