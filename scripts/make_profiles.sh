#!/bin/bash
# Round profile artefacts (run under gpurun, one GPU): ncu launch list of a short RL run and a full
# capture of one RL iteration's kernels.  Raw reports land in gpurun_out/, summaries in profiles/.
cd "$(dirname "$0")/.."
TAG=${1:-r02}
export PROBE_SHAPE=${PROBE_SHAPE:-512,512,512}   # the metric's own configuration
mkdir -p gpurun_out profiles
export PROBE_ITERS=3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/prof_run.py > gpurun_out/launches_$TAG.log 2>&1
export PROBE_ITERS=2
ncu --set full --clock-control none --import-source on -k regex:'k_ypassT|k_zconvT|k_zrow|k_ypassF|k_xpassP' -s 6 -c 8 -o gpurun_out/prof_$TAG -f python scripts/prof_run.py > gpurun_out/prof_full_$TAG.log 2>&1
# registration cost kernel
cat > /tmp/prof_reg.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from microimagelib_b200 import device, synth
shape = (256, 512, 512)
psf = synth.gaussian_psf((33, 33, 33), (4, 2, 2))
img = synth.bead_image(shape, psf, noise=False)
m = synth.affine_matrix(2.0, (1.02, 0.99, 1.0), (3.5, -2.25, 1.75), center=(256, 256, 128))
t = torch.from_numpy(img).cuda()
r = device.Reg(shape); r.set_images(t, t); r.prepare()
for _ in range(3): r.cost(m)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:'^k_zncc$' --kernel-name-base function -s 2 -c 1 -o gpurun_out/prof_zncc_$TAG -f python /tmp/prof_reg.py > gpurun_out/prof_zncc_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
