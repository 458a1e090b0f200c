#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- X tiles at N=512: default vs wide"
AB_SHAPE=512,512,512 AB_TAG="default 512^3" timeout 300 python scripts/ab_iter.py
AB_SHAPE=512,512,512 AB_TAG="xwide512 512^3" MILB_LIBAPI=$PWD/microimagelib_b200/lib_alt/libapi_xwide512.so timeout 300 python scripts/ab_iter.py
echo "--- reference diagnostics"
python scripts/ref_diag.py 2>&1 | grep -v "^\.\.\.\|^Image\|^GPU\|^$" | tail -16
echo "--- gpu tests"
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/gpu_tests6.log; tail -12 gpurun_out/gpu_tests6.log
cat > /tmp/zk.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from microimagelib_b200 import device
shape = (256, 512, 512)
vol = torch.rand(shape, device="cuda") * 100 + 10
m = np.array([0.9994, 0.0349, 0, -5.1, -0.0349, 0.9994, 0, 6.3, 0, 0, 1, 1.75], np.float32)
r = device.Reg(shape); r.set_images(vol, vol); r.prepare()
for _ in range(3): r.cost(m)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:'^k_zncc$' --kernel-name-base function -s 1 -c 1 -o gpurun_out/zncc_hw_k1 -f python /tmp/zk.py > gpurun_out/ncu_zncc.log 2>&1; tail -2 gpurun_out/ncu_zncc.log
