import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microimagelib_b200 import _lib
lib = _lib.load()
F = C.POINTER(C.c_float)
lib.milb_debug_tex3d_sample.argtypes = [F, F, C.POINTER(C.c_uint), F, C.c_int, C.c_int]
rng = np.random.default_rng(0)
vol = (rng.random((16, 16, 16)) * 1000).astype(np.float32)
size = (C.c_uint * 3)(16, 16, 16)
out = {}
n = 20000
for name, axes in (("x", [0]), ("y", [1]), ("z", [2]), ("xy", [0, 1]), ("xz", [0, 2]), ("yz", [1, 2]), ("xyz", [0, 1, 2])):
    c = np.floor(rng.random((n, 3)) * 14 + 1) + 0.5          # texel centres -> alpha = 0
    for a in axes:
        c[:, a] = rng.random(n) * 14 + 1.0
    c = c.astype(np.float32)
    r = {}
    for hw in (1, 0):
        o = np.zeros(n, np.float32)
        lib.milb_debug_tex3d_sample(o.ctypes.data_as(F), vol.ctypes.data_as(F), size, c.ctypes.data_as(F), n, hw)
        r[hw] = o
    e = np.abs(r[1] - r[0])
    print(name, "max", e.max(), "mean", e.mean())
    out["c_" + name] = c; out["hw_" + name] = r[1]; out["sw_" + name] = r[0]
np.savez(os.path.join(ROOT, "gpurun_out", "tex_probe3.npz"), vol=vol, **out)
