import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microimagelib_b200 import device, synth
shape = tuple(int(v) for v in sys.argv[1].split(","))
psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
img = np.random.default_rng(1).random(shape, dtype=np.float32) + 0.1
d = device.Decon(shape, 1, row_conv=True)
d.set_psf(0, psf); d.set_image(0, img)
st = torch.cuda.current_stream()
d.run(2, stream=st); torch.cuda.synchronize(); print("run ok", flush=True)
print(d.time_kernels(1, st).tolist(), flush=True)
torch.cuda.synchronize(); print("time_kernels ok", flush=True)
