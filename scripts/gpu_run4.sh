#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- L2-resident per-tile probe (17 planes of 512x512: both spectra fit in L2; 544 tiles = 4 per CTA at most)"
AB_SHAPE=30,512,512 AB_TAG="unfused 32x512x512" timeout 300 python scripts/ab_iter.py
AB_SHAPE=62,512,512 AB_TAG="unfused 64x512x512" timeout 300 python scripts/ab_iter.py
echo "--- zncc"
python scripts/zncc_probe.py 2>&1 | tail -14
echo "--- reference diagnostics"
python scripts/ref_diag.py 2>&1 | grep -v "^\.\.\.\|^Image\|^GPU\|^$" | tail -40
