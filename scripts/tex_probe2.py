"""Reveals the hardware's texture weight quantisation: sample a ramp at fine sub-voxel steps."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microimagelib_b200 import _lib
lib = _lib.load()
F = C.POINTER(C.c_float)
lib.milb_debug_tex3d_sample.argtypes = [F, F, C.POINTER(C.c_uint), F, C.c_int, C.c_int]
src = np.zeros((8, 8, 8), np.float32)
src[:] = np.arange(8, dtype=np.float32)[None, None, :]       # ramp along x
n = 8192
sub = np.arange(n, dtype=np.float64) / 4096.0                   # two voxels, 4096 steps each
coords = np.zeros((n, 3), np.float32)
coords[:, 0] = (2.5 + sub).astype(np.float32)                  # xb = 2.0 .. 4.0
coords[:, 1] = 3.5
coords[:, 2] = 3.5
size = (C.c_uint * 3)(8, 8, 8)
res = {}
for hw in (1, 0):
    out = np.zeros(n, np.float32)
    assert lib.milb_debug_tex3d_sample(out.ctypes.data_as(F), src.ctypes.data_as(F), size, coords.ctypes.data_as(F), n, hw) == 0
    res[hw] = out
# random 3-D samples on a random volume: hw vs sw
rng = np.random.default_rng(0)
vol = (rng.random((16, 16, 16)) * 1000).astype(np.float32)
c2 = (rng.random((20000, 3)) * 15 + 0.5).astype(np.float32)
size2 = (C.c_uint * 3)(16, 16, 16)
for hw in (1, 0):
    out = np.zeros(len(c2), np.float32)
    lib.milb_debug_tex3d_sample(out.ctypes.data_as(F), vol.ctypes.data_as(F), size2, c2.ctypes.data_as(F), len(c2), hw)
    res[f"r{hw}"] = out
np.savez(os.path.join(ROOT, "gpurun_out", "tex_probe2.npz"), coords=coords, hw=res[1], sw=res[0], vol=vol, c2=c2, rhw=res["r1"], rsw=res["r0"])
print("ramp max |hw-sw|", np.abs(res[1] - res[0]).max(), "random: max", np.abs(res["r1"] - res["r0"]).max(), "mean", np.abs(res["r1"] - res["r0"]).mean())
