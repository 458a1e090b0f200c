"""Row convolution (k_zrow) against the transposing plane kernels (k_ypassT / k_zconvT): result agreement on boxes that
cover every two-stage Z length, then per-kernel and per-iteration times at the bench boxes.  One JSON line per item."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from microimagelib_b200 import device, synth


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))


def run(shape, row, iters, dual=False):
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
    img = synth.bead_image(shape, psf, density=1 / 4096.0)
    d = device.Decon(shape, 2 if dual else 1, row_conv=row)
    for v in range(2 if dual else 1):
        d.set_psf(v, psf)
        d.set_image(v, img)
    d.run(iters)
    out = d.result().copy()
    d.close()
    return out


def timing(shape, row, iters=20):
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
    rng = np.random.default_rng(1)
    img = (rng.random(shape, dtype=np.float32) + 0.1)
    d = device.Decon(shape, 1, row_conv=row)
    d.set_psf(0, psf)
    d.set_image(0, img)
    st = torch.cuda.current_stream()
    d.run(3, stream=st)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        d.run(iters, stream=st)
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    k = d.time_kernels(5, st).tolist()
    d.close()
    n = float(np.prod(shape))
    return {"box": list(shape), "row_conv": row, "ms_per_iteration": best, "frac_of_6552": 56.0 * n / (best * 1e-3) / 6552e9, "kernels_ms": k}


if __name__ == "__main__":
    torch.cuda.init()
    for shape in [(64, 64, 64), (64, 128, 128), (128, 64, 256), (64, 192, 512), (256, 256, 256), (64, 64, 1024), (128, 320, 128)]:
        a, b = run(shape, True, 5), run(shape, False, 5)
        print(json.dumps({"box": list(shape), "rel_l2_row_vs_transposing": rel_l2(a, b), "max": float(np.abs(a - b).max())}), flush=True)
    a, b = run((64, 128, 256), True, 4, dual=True), run((64, 128, 256), False, 4, dual=True)
    print(json.dumps({"box": [64, 128, 256], "dual": True, "rel_l2_row_vs_transposing": rel_l2(a, b)}), flush=True)
    for shape in [(256, 512, 512), (512, 512, 512), (256, 256, 256), (128, 1024, 1024)]:
        for row in (False, True):
            print(json.dumps(timing(shape, row)), flush=True)
