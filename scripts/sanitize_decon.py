"""Small single-view runs for compute-sanitizer (memcheck / racecheck / synccheck): row convolution at Z = 512 and 256, folded X pass
at X = 512 and 256, plane pipeline (PIPE=1 in argv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if len(sys.argv) > 2 and sys.argv[2] == "pipe":
    os.environ["MILB_PLANE_PIPE"] = "1"
from microimagelib_b200 import device, synth
shape = tuple(int(v) for v in sys.argv[1].split(","))
psf = synth.gaussian_psf((9, 9, 9), (1.5, 1.5, 1.5))
img = np.random.default_rng(1).random(shape, dtype=np.float32) + 0.1
d = device.Decon(shape, 1)
d.set_psf(0, psf); d.set_image(0, img)
d.run(2)
out = d.result()
print("ok", shape, float(out.sum()))
