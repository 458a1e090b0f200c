"""Per-kernel times of the plane kernels on a SUBSET of the SMs (MILB_GRID_CAP): with HBM no longer the limit, time x CTAs / tiles
is what a tile costs the SM itself -- the quantity that decides whether running the kernels side by side (plane pipeline) can pay."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microimagelib_b200 import device, synth
shape = tuple(int(v) for v in os.environ.get("PROBE_SHAPE", "256,512,512").split(","))
psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
img = np.random.default_rng(1).random(shape, dtype=np.float32) + 0.1
d = device.Decon(shape, 1)
d.set_psf(0, psf); d.set_image(0, img)
st = torch.cuda.current_stream()
d.run(2, stream=st); torch.cuda.synchronize()
k = d.time_kernels(5, st).tolist()
print(json.dumps({"cap": os.environ.get("MILB_GRID_CAP", "0"), "box": list(shape), "yfwd_zrow_yinv_ms": k[:3]}))
