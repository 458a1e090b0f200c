"""Diagnostics of product / oracle against the reference's own GPU build (oracle/_ref): where the warp differs, and how
the whole registration (records, matrices) compares per affMethod for reference / product hw fetch / product sw fetch / oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from microimagelib_b200 import synth, libapi, device, _lib
from oracle import ref_gpu, reg_oracle as ro

IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)

def corner_disp(m1, m2, shape):
    sz, sy, sx = shape
    d = (np.asarray(m1, np.float64) - np.asarray(m2, np.float64)).reshape(3, 4)
    return max(float(np.linalg.norm(d @ np.array([x, y, z, 1.0]))) for x in (0, sx - 1) for y in (0, sy - 1) for z in (0, sz - 1))

def pair(shape=(40, 56, 72), seed=3):
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 2.0))
    tgt = synth.bead_image(shape, psf, seed=seed, density=1 / 2048.0)
    sz, sy, sx = shape
    m = synth.affine_matrix(rot_z_deg=2.0, scale=(1.02, 0.99, 1.0), shift=(1.5, -1.25, 0.75), center=(sx / 2, sy / 2, sz / 2))
    return tgt, synth.warp_exact(tgt, m), m

R = ref_gpu.api()
tgt, src, m = pair()
lib = _lib.load()
F = C.POINTER(C.c_float)
lib.milb_debug_tex3d_warp.argtypes = [F, F, C.POINTER(C.c_uint), F]
print("== warp: reference (tex3D) vs oracle (software) vs hardware fetch with the product's coordinates")
big = synth.affine_matrix(rot_z_deg=35.0, scale=(0.7, 1.3, 1.0), shift=(4, -3, 2))
for name, mat in (("2deg", m), ("ident", IDENT), ("35deg", big)):
    ref, _ = R.atrans3dgpu(src, mat)
    orc = ro.affine_warp(src, mat)
    hw = np.zeros_like(src)
    size = (C.c_uint * 3)(src.shape[2], src.shape[1], src.shape[0])
    mm = np.ascontiguousarray(mat, np.float32)
    lib.milb_debug_tex3d_warp(hw.ctypes.data_as(F), src.ctypes.data_as(F), size, mm.ctypes.data_as(F))
    scale = float(np.abs(src).max())
    d_ro, d_rh = np.abs(ref - orc), np.abs(ref - hw)
    print(f"  {name}: max|ref-orc| {d_ro.max():.4g} ({d_ro.max()/scale:.2e} rel)  n>1e-5*scale: {(d_ro > 1e-5*scale).sum()}  mask diff: {((ref == 0) != (orc == 0)).sum()}"
          f" | max|ref-hw(my coords)| {d_rh.max():.4g}  n!=: {(ref != hw).sum()} of {ref.size}")
    if d_ro.max() > 1e-5 * scale:
        idx = np.unravel_index(np.argmax(d_ro), d_ro.shape)
        print("    worst voxel (z,y,x)", idx, "ref", ref[idx], "orc", orc[idx], "hw", hw[idx])

print("== reg3d per affMethod: [initial ZNCC, final ZNCC, evals] and matrix displacement vs the reference")
for method in (1, 2, 3, 4, 5, 6, 7):
    _, tr, _, rr = R.reg3d(tgt, src, regChoice=2, regMethod=method, FTOL=1e-4, itLimit=3000)
    os.environ["MILB_ZNCC_FETCH"] = "hw"
    _, th, _, rh = libapi.reg3d(tgt, src, regChoice=2, regMethod=method, FTOL=1e-4, itLimit=3000)
    os.environ["MILB_ZNCC_FETCH"] = "sw"
    _, ts, _, rs = libapi.reg3d(tgt, src, regChoice=2, regMethod=method, FTOL=1e-4, itLimit=3000)
    os.environ.pop("MILB_ZNCC_FETCH")
    o = ro.reg3d_affine(tgt, src, method, ftol=1e-4, it_limit=3000)
    def f(r): return f"[{float(r[1]):.7f} {float(r[3]):.7f} {int(r[5])}]"
    print(f"  m{method}: ref {f(rr)} hw {f(rh)} sw {f(rs)} orc {f(o['records'])} | disp hw {corner_disp(th, tr, tgt.shape):.2e} sw {corner_disp(ts, tr, tgt.shape):.2e} orc {corner_disp(o['tmx'], tr, tgt.shape):.2e}")
