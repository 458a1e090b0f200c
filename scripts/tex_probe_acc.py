#!/usr/bin/env python
"""Collects hardware tex2D / tex3D outputs on random data so that the unit's final accumulation
(order / precision of the weighted sum) can be studied offline: gpurun_out/tex_acc.npz."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from microimagelib_b200 import _lib, device  # noqa: E402

rng = np.random.default_rng(7)
img = (rng.random((37, 53)) * 1000).astype(np.float32)
n = 50000
c2 = np.stack([rng.uniform(0.5, 52.5, n), rng.uniform(0.5, 36.5, n)], 1).astype(np.float32)
hw2 = device.tex2d_samples(img, c2, hardware=True)
sw2 = device.tex2d_samples(img, c2, hardware=False)
lib = _lib.load()
F = C.POINTER(C.c_float)
lib.milb_debug_tex3d_sample.argtypes = [F, F, C.POINTER(C.c_uint), F, C.c_int, C.c_int]
vol = (rng.random((16, 20, 24)) * 1000).astype(np.float32)
c3 = np.stack([rng.random(n) * 23 + 0.5, rng.random(n) * 19 + 0.5, rng.random(n) * 15 + 0.5], axis=1).astype(np.float32)
size = (C.c_uint * 3)(24, 20, 16)
out = {}
for hw in (1, 0):
    o = np.zeros(n, np.float32)
    assert lib.milb_debug_tex3d_sample(o.ctypes.data_as(F), vol.ctypes.data_as(F), size, c3.ctypes.data_as(F), n, hw) == 0
    out[hw] = o
print("2-D: differ", np.count_nonzero(hw2 != sw2), "3-D: differ", np.count_nonzero(out[1] != out[0]), "of", n)
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/tex_acc.npz", img=img, c2=c2, hw2=hw2, sw2=sw2, vol=vol, c3=c3, hw3=out[1], sw3=out[0])
