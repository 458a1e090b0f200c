"""CPU side of scripts/tex_cases.py: where does the software restatement of the texture filter (oracle/reg_oracle.c)
differ from the reference's hardware fetches?  Prints the mismatching samples with their coordinates and fractions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import reg_oracle as ro
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tex_cases import cases

def coords(m, shape):
    """aff_coord in float32 with the reference build's contraction (a1*y, fma a0*x, fma a2*z, + a3, + 0.5)"""
    sz, sy, sx = shape
    z, y, x = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    out = []
    for r in range(3):
        a = m[4 * r: 4 * r + 4].astype(np.float32)
        t = (a[1] * y.astype(np.float32)).astype(np.float32)
        t = (np.float64(a[0]) * x + np.float64(t)).astype(np.float32)      # fma: exact product + t, one rounding (float64 holds it exactly enough)
        t = (np.float64(a[2]) * z + np.float64(t)).astype(np.float32)
        t = (t + a[3]).astype(np.float32)
        t = (t + np.float32(0.5)).astype(np.float32)
        out.append(t)
    return out

vol, mats = cases()
d = np.load("gpurun_out/tex_cases.npz")
tot = 0
for k, m in enumerate(mats):
    ref = d["outs"][k]
    orc = ro.affine_warp(vol, m)
    diff = np.abs(ref - orc)
    bad = np.argwhere(diff > 0.004)           # values up to 1000: anything above float rounding of the sum
    tx, ty, tz = coords(m, vol.shape)
    print(f"case {k}: {len(bad)} of {ref.size} samples differ by more than rounding; max {diff.max():.4f}")
    tot += len(bad)
    for (z, y, x) in bad[:12]:
        c = [float(t[z, y, x]) for t in (tx, ty, tz)]
        fr = [(v - 0.5) - np.floor(v - 0.5) for v in c]
        print(f"   vox ({x},{y},{z}) coord {c[0]:.6f} {c[1]:.6f} {c[2]:.6f}  frac*256 {fr[0]*256:.4f} {fr[1]*256:.4f} {fr[2]*256:.4f}  ref {ref[z,y,x]:.4f} orc {orc[z,y,x]:.4f}")
print("total", tot)
