"""Plane pipeline (Y forward | row convolution | Y inverse side by side, MILB_PLANE_PIPE=1) against the same kernels one after
the other: bit-identity on small boxes, then per-iteration time at the bench boxes for a few SM splits."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from microimagelib_b200 import device, synth


def make(shape, pipe, split=None):
    os.environ["MILB_PLANE_PIPE"] = "1" if pipe else "0"
    if split:
        os.environ["MILB_PIPE_SPLIT"] = split
    else:
        os.environ.pop("MILB_PIPE_SPLIT", None)
    return device.Decon(shape, 1)


def run(shape, pipe, iters=5):
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
    img = synth.bead_image(shape, psf, density=1 / 4096.0)
    d = make(shape, pipe)
    d.set_psf(0, psf)
    d.set_image(0, img)
    d.run(iters)
    out = d.result().copy()
    d.close()
    return out


def timing(shape, pipe, split=None, iters=20):
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
    img = np.random.default_rng(1).random(shape, dtype=np.float32) + 0.1
    d = make(shape, pipe, split)
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
    roles = d.time_pipe(6, st).tolist() if pipe else None
    d.close()
    n = float(np.prod(shape))
    return {"box": list(shape), "pipe": pipe, "split": split, "ms_per_iteration": best, "frac_of_6552": 56.0 * n / (best * 1e-3) / 6552e9,
            "yfwd_zrow_yinv_stage_ms": roles}


if __name__ == "__main__":
    torch.cuda.init()
    for shape in [(64, 256, 256), (128, 512, 512)]:
        a, b = run(shape, True), run(shape, False)
        print(json.dumps({"box": list(shape), "pipe_equals_sequential": bool(np.array_equal(a, b)), "max_abs_diff": float(np.abs(a - b).max())}), flush=True)
    splits = sys.argv[1:] or ["0.27,0.46", "0.24,0.52", "0.20,0.60"]
    for shape in [(256, 512, 512), (512, 512, 512)]:
        print(json.dumps(timing(shape, False)), flush=True)
        for s in splits:
            print(json.dumps(timing(shape, True, s)), flush=True)
