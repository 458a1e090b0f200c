#!/bin/bash
# Does running the plane passes chunk by chunk keep the intermediates in L2?  DRAM bytes, L2 hit rate and
# duration per launch for chunk = all / 8 / 16 planes -> gpurun_out/l2_chunk<N>.csv
cd "$(dirname "$0")/.."
for CH in 0 8 16; do
  export PROBE_ITERS=2 PROBE_CHUNK=$CH
  SKIP=$(( CH == 0 ? 10 : (CH == 8 ? 120 : 70) ))
  CNT=$(( CH == 0 ? 16 : (CH == 8 ? 110 : 60) ))
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
      -k regex:'k_ypassT|k_zconvT|k_ypassF|k_xpassP' -s $SKIP -c $CNT --csv --log-file gpurun_out/l2_chunk$CH.csv \
      python scripts/prof_run.py > gpurun_out/l2_chunk$CH.log 2>&1
  tail -1 gpurun_out/l2_chunk$CH.log
done
for CH in 0 8 16 32; do
  PROBE_CHUNK=$CH PROBE_ITERS=30 python - <<'PY'
import os, sys, time
sys.path.insert(0, '.')
import torch
from microimagelib_b200 import device, synth
shape = (256, 512, 512)
psf = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
d = device.Decon(shape, 1)
d.set_psf(0, psf)
d.set_image(0, torch.rand(shape, device='cuda') * 100 + 10)
d.set_chunk_planes(int(os.environ['PROBE_CHUNK']))
d.run(3)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); d.run(30); b.record(); torch.cuda.synchronize()
print('chunk', os.environ['PROBE_CHUNK'], 'ms/iter', a.elapsed_time(b) / 30)
PY
done
