#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- zncc"; timeout 300 python scripts/zncc_probe.py 2>&1 | tail -8
echo "--- config 4"; timeout 300 python scripts/config4_reg.py 2>&1 | tail -1
echo "--- gpu tests"; python -m pytest tests -q -m gpu 2>&1 | tail -5
