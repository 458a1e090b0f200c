#!/usr/bin/env python
"""Where does the end-to-end call (libapi.decon_singleview, host buffers) spend its time beyond the loop?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microimagelib_b200 import device, libapi, synth

shape = (256, 512, 512)
iters = 50
psf = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
img = (np.random.default_rng(0).random(shape, dtype=np.float32) * 100 + 10)
h_img = torch.from_numpy(img).pin_memory().numpy()
h_out = torch.empty(shape, dtype=torch.float32).pin_memory().numpy()

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

devnull = os.open(os.devnull, os.O_WRONLY)
print("libapi.decon_singleview pinned   : %.2f ms" % t(lambda: libapi.decon_singleview(h_img, psf, iters, out=h_out)))
print("libapi.decon_singleview pageable : %.2f ms" % t(lambda: libapi.decon_singleview(img, psf, iters)))
d = device.Decon(shape, 1)
d.set_psf(0, psf)
print("set_image (H2D pinned + pad)     : %.2f ms" % t(lambda: d.set_image(0, h_img)))
d_img = torch.from_numpy(img).cuda()
print("set_image (device + pad)         : %.2f ms" % t(lambda: d.set_image(0, d_img)))
print("run(%d)                          : %.2f ms" % (iters, t(lambda: d.run(iters))))
print("run(0)                           : %.2f ms" % t(lambda: d.run(0)))
print("result (crop + D2H pinned)       : %.2f ms" % t(lambda: d.result(out=h_out)))
x = torch.empty(shape, dtype=torch.float32, device="cuda")
hp = torch.from_numpy(h_img)
print("torch H2D pinned 268 MB          : %.2f ms" % t(lambda: x.copy_(hp, non_blocking=True)))
ho = torch.from_numpy(h_out)
print("torch D2H pinned 268 MB          : %.2f ms" % t(lambda: ho.copy_(x, non_blocking=True)))
print("cudaMemGetInfo x4                : %.2f ms" % t(lambda: [torch.cuda.mem_get_info() for _ in range(4)]))
