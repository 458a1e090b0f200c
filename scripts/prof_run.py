"""Small driver for profiling: a few RL iterations on the bench volume."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from microimagelib_b200 import device, synth

shape = tuple(int(x) for x in os.environ.get("PROBE_SHAPE", "256,512,512").split(","))
psf = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
img = synth.bead_image(shape, psf, noise=False)  # noise-free: faster to generate, same kernels
d = device.Decon(shape, 1)
d.set_psf(0, psf)
d.set_image(0, torch.from_numpy(img).cuda())
ch = os.environ.get("PROBE_CHUNK")
if ch is not None:
    d.set_chunk_planes(int(ch))
d.run(int(os.environ.get("PROBE_ITERS", "3")))
torch.cuda.synchronize()
