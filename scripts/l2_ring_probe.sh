#!/bin/bash
# Round 2: do the transposed planes stay in L2 when they live in a small ring of S2 slots (MILB_RING_PLANES)
# and the three plane passes run chunk by chunk?  DRAM bytes + duration per launch (ncu) and ms/iteration.
cd "$(dirname "$0")/.."
for CFG in "0 0" "8 0" "8 32" "4 16" "16 48"; do
  set -- $CFG
  export PROBE_ITERS=2 PROBE_CHUNK=$1 MILB_RING_PLANES=$2
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
      -k regex:'k_ypassT|k_zconvT|k_ypassF|k_xpassP' --csv --log-file gpurun_out/ring_c$1_r$2.csv \
      python scripts/prof_run.py > gpurun_out/ring_c$1_r$2.log 2>&1
  tail -1 gpurun_out/ring_c$1_r$2.log
done
for CFG in "0 0" "8 0" "8 32" "4 16" "16 48" "2 8"; do
  set -- $CFG
  PROBE_CHUNK=$1 MILB_RING_PLANES=$2 python - <<'PY'
import os, sys
sys.path.insert(0, '.')
import torch
from microimagelib_b200 import device, synth
shape = (256, 512, 512)
psf = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
d = device.Decon(shape, 1)
d.set_psf(0, psf)
g = torch.Generator(device="cuda"); g.manual_seed(1)
d.set_image(0, torch.rand(shape, device='cuda', generator=g) * 100 + 10)
d.set_chunk_planes(int(os.environ['PROBE_CHUNK']))
d.run(3)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); d.run(30); b.record(); torch.cuda.synchronize()
d.run(7)
print('chunk', os.environ['PROBE_CHUNK'], 'ring', os.environ['MILB_RING_PLANES'], 'ms/iter', a.elapsed_time(b) / 30, 'checksum', float(torch.as_tensor(d.result()).double().sum()))
PY
done
