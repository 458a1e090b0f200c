#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/zrow_repro.py 256,512,512 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/zrow_repro.py 64,128,128 2>&1 | grep -v "^=========     Host Frame\|^=========         in " | head -60
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/zrow_repro.py 256,512,512 2>&1 | grep -v "^=========     Host Frame\|^=========         in " | head -60
