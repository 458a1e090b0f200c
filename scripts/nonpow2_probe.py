"""ms per RL iteration on boxes snapTransformSize produces for non-power-of-two images: the compile-time 64*k plans (fast) against
the generic mixed-radix kernels (MILB_FORCE_GENERIC=1), with the same roofline record as bench.py (56 B per voxel and iteration)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch
from microimagelib_b200 import device, synth
shape = tuple(int(v) for v in os.environ["NP_SHAPE"].split(","))
psf = synth.gaussian_psf((33, 33, 33), (4, 2, 2))
d = device.Decon(shape, 1)
d.set_psf(0, psf)
g = torch.Generator(device="cuda"); g.manual_seed(1)
d.set_image(0, torch.rand(shape, device="cuda", generator=g) * 100 + 10)
d.run(5); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); d.run(20); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 20)
d.run(8)
n = 1
for s in d.fft_shape: n *= s
print(json.dumps({"box": list(d.fft_shape), "ms_per_iteration": best, "GBps": 56 * n / best / 1e6, "frac_of_6552": 56 * n / best / 1e6 / 6552,
                  "checksum": float(torch.as_tensor(d.result()).double().sum())}))
''' % ROOT
out = {}
for shape in ("192,320,320", "320,576,576", "256,448,448", "384,640,640", "256,512,512"):
    row = {}
    for tag, env in (("fast", "0"), ("generic", "1")):
        r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env={**os.environ, "NP_SHAPE": shape, "MILB_FORCE_GENERIC": env})
        try:
            row[tag] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            row[tag] = {"error": (r.stderr or r.stdout)[-300:]}
    if "ms_per_iteration" in row.get("fast", {}) and "ms_per_iteration" in row.get("generic", {}):
        row["speedup_fast_vs_generic"] = row["generic"]["ms_per_iteration"] / row["fast"]["ms_per_iteration"]
        row["rel_checksum_diff"] = abs(row["fast"]["checksum"] - row["generic"]["checksum"]) / abs(row["generic"]["checksum"])
    out[shape] = row
    print(shape, json.dumps(row), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "nonpow2_probe.json"), "w"), indent=1)
