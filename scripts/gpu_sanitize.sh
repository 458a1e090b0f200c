#!/bin/bash
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck; do
  for cfg in "512,64,512" "256,64,256" "64,512,512 pipe"; do
    echo "--- $tool $cfg"
    timeout 600 compute-sanitizer --tool $tool --print-limit 3 python scripts/sanitize_decon.py $cfg 2>&1 | grep -v "Host Frame\|^=========         in \|^=========     Saved host" | grep "ERROR SUMMARY\|Invalid\|hazard\|Race\|ok \|Error\|error" | head -8
  done
done
