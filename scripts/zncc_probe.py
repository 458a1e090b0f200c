"""ms per ZNCC evaluation at BASELINE config 4's size for the hardware-texture and software fetch paths (K = 1, 4, 8
candidates per launch), and the reference's own corrkernel on the same GPU (oracle/_ref) for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from microimagelib_b200 import device, synth

shape = (256, 512, 512)
g = torch.Generator(device="cuda"); g.manual_seed(1)
vol = torch.rand(shape, device="cuda", generator=g) * 100 + 10
m = np.array([0.9994, 0.0349, 0, -5.1, -0.0349, 0.9994, 0, 6.3, 0, 0, 1, 1.75], np.float32)
big = synth.affine_matrix(rot_z_deg=35.0, scale=(0.9, 1.1, 1.0), shift=(4, -3, 2), center=(256, 256, 128))
n = float(np.prod(shape))
for mode in ("hw", "sw"):
    r = device.Reg(shape, fetch=mode)
    r.set_images(vol, vol)
    r.prepare()
    for name, mat in (("2deg", m), ("35deg", big)):
        for K in (1, 4, 8):
            mats = np.stack([mat] * K); mats[:, 3] += 0.1 * np.arange(K, dtype=np.float32)
            r.cost(mats); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): r.cost(mats)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10 / K
            print(f"{mode} {name} K={K}: {ms:.4f} ms/eval  {8*n/ms/1e6:.0f} GB/s  frac {8*n/ms/1e6/6552:.3f}", flush=True)
    r.close()
try:
    from oracle import ref_gpu
    if ref_gpu.available():
        h = vol.cpu().numpy()
        _, _, st, rec = ref_gpu.api().reg3d(h, h, regChoice=2, regMethod=5, inputTmx=True, iTmx=m, itLimit=1)
        print("reference corrfunc (its records[4], ms per evaluation incl. host):", float(rec[4]), "evals", float(rec[5]), "iter s", float(rec[6]))
except Exception as e:
    print("reference unavailable:", e)
