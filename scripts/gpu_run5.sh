#!/bin/bash
cd "$(dirname "$0")/.."
python tests/golden/make_reference_golden.py gpurun_out/reference_vectors.npz 2>&1 | tail -2
python scripts/tex_cases.py 2>&1 | tail -1
cp gpurun_out/reference_vectors.npz tests/golden/reference_vectors.npz
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/gpu_tests5.log; tail -30 gpurun_out/gpu_tests5.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -c 3000 gpurun_out/bench5.json; tail -5 gpurun_out/bench5.err
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
ncu --set full --clock-control none --import-source on -k regex:k_zncc -s 1 -c 1 -o gpurun_out/zncc_hw_k1 python /tmp/zk.py > gpurun_out/ncu_zncc.log 2>&1; tail -2 gpurun_out/ncu_zncc.log
