"""BASELINE config 4: 512x512x256 target, source = target warped by 2 deg about z, scale (1.02, 0.99, 1.0),
shift (3.5, -2.25, 1.75); reg3d(regChoice=2, affMethod=7, FTOL=1e-4, itLimit=3000).  Prints wall time,
evaluations, ZNCC and the recovered matrix's worst corner displacement from the inverse of the applied one."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from microimagelib_b200 import device, libapi, synth

shape = (256, 512, 512)
psf = synth.gaussian_psf((33, 33, 33), (4, 2, 2))
tgt = synth.bead_image(shape, psf, seed=20260)
m = synth.affine_matrix(2.0, (1.02, 0.99, 1.0), (3.5, -2.25, 1.75), center=(shape[2] / 2, shape[1] / 2, shape[0] / 2))
src = device.affine_warp(tgt, m)                      # src(x) = tgt(M x)
out = {}
for method in (7, 6):
    t0 = time.perf_counter()
    reg, tmx, st, rec = libapi.reg3d(tgt, src, regChoice=2, regMethod=method, FTOL=1e-4, itLimit=3000)
    dt = time.perf_counter() - t0
    inv = synth.invert_affine(m)
    d = (np.asarray(tmx, np.float64) - np.asarray(inv, np.float64)).reshape(3, 4)
    worst = max(float(np.linalg.norm(d @ np.array([x, y, z, 1.0]))) for x in (0, shape[2] - 1) for y in (0, shape[1] - 1) for z in (0, shape[0] - 1))
    out[f"affMethod{method}"] = {"status": int(st), "wall_s": dt, "evaluations": int(rec[5]), "zncc_initial": float(rec[1]), "zncc_final": float(rec[3]),
                                 "ms_per_evaluation_incl_host": 1e3 * dt / max(int(rec[5]), 1), "worst_corner_displacement_vs_truth_voxels": worst}
print(json.dumps(out))
