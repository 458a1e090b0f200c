import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microimagelib_b200 import libapi, synth
shape = (256, 512, 512)
psf = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
img = (np.random.default_rng(0).random(shape, dtype=np.float32) * 100 + 10)
h_img = torch.from_numpy(img).pin_memory().numpy()
h_out = torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
for i in range(5):
    t0 = time.perf_counter()
    out, st, rec = libapi.decon_singleview(h_img, psf, 50, out=h_out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    sys.stderr.write("call %d: %.2f ms  records init %.2f prep %.2f run+d2h %.2f total %.2f ms\n" % (i, dt, rec[6]*1e3, rec[7]*1e3, rec[8]*1e3, rec[9]*1e3))
