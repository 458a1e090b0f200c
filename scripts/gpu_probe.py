"""Ad-hoc GPU measurements for development (run under gpurun); writes gpurun_out/probe.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microimagelib_b200 import _lib, device, synth  # noqa: E402

out = {}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
what = sys.argv[1:] or ["tex", "decon", "reg"]


def ev_time(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if "tex" in what:
    lib = _lib.load()
    F = C.POINTER(C.c_float)
    lib.milb_debug_tex3d_warp.argtypes = [F, F, C.POINTER(C.c_uint), F]
    rng = np.random.default_rng(0)
    src = (rng.random((24, 28, 32)) * 1000).astype(np.float32)
    mats = [synth.affine_matrix(3.0, (1.03, 0.97, 1.01), (0.37, -0.61, 0.22)), synth.affine_matrix(-11.0, (0.9, 1.1, 1.0), (2.3, 1.7, -0.9))]
    hw = []
    for m in mats:
        o = np.zeros_like(src)
        size = (C.c_uint * 3)(src.shape[2], src.shape[1], src.shape[0])
        lib.milb_debug_tex3d_warp(o.ctypes.data_as(F), src.ctypes.data_as(F), size, m.ctypes.data_as(F))
        hw.append(o)
    np.savez(os.path.join(ROOT, "gpurun_out", "tex_probe.npz"), src=src, mats=np.stack(mats), hw=np.stack(hw))

if "decon" in what:
    shape = tuple(int(x) for x in os.environ.get("PROBE_SHAPE", "256,512,512").split(","))
    psf = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
    img = synth.bead_image(shape, psf)
    d = device.Decon(shape, 1)
    d.set_psf(0, psf)
    d.set_image(0, torch.from_numpy(img).cuda())
    res = {}
    for chunk in (0, 8, 16, 32, 64):
        d.set_chunk_planes(chunk)
        d.run(2)
        res[f"chunk{chunk}_ms_per_iter"] = ev_time(lambda: d.run(10)) / 10
    d.set_chunk_planes(0)
    d.run(10)
    mine = d.result().copy()
    res["yardstick_ms_per_iter"] = d.run_cufft_yardstick(10) / 10
    yard = d.result()
    res["rel_l2_vs_cufft_yardstick_10it"] = float(np.linalg.norm(mine.astype(np.float64) - yard) / np.linalg.norm(yard.astype(np.float64)))
    out["decon"] = res
    d.close()

if "reg" in what:
    shape = tuple(int(x) for x in os.environ.get("PROBE_SHAPE", "256,512,512").split(","))
    psf = synth.gaussian_psf((33, 33, 33), (4, 2, 2))
    img = synth.bead_image(shape, psf)
    m = synth.affine_matrix(2.0, (1.02, 0.99, 1.0), (3.5, -2.25, 1.75), center=(shape[2] / 2, shape[1] / 2, shape[0] / 2))
    t = torch.from_numpy(img).cuda()
    r = device.Reg(shape)
    r.set_images(t, t)
    r.prepare()
    res = {}
    for K in (1, 2, 4, 8):
        mats = np.stack([m] * K)
        mats[:, 3] += np.arange(K) * 0.1
        r.cost(mats)
        t0 = time.perf_counter()
        for _ in range(5):
            r.cost(mats)
        res[f"K{K}_ms_per_launch"] = (time.perf_counter() - t0) / 5 * 1e3
    out["reg"] = res
    r.close()

print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
