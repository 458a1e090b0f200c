"""ms per RL iteration at 512x512x256 (single view) + parity of the result against a checksum; A/B helper"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from microimagelib_b200 import device, synth
shape = tuple(int(v) for v in os.environ.get("AB_SHAPE", "256,512,512").split(","))
psf = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
d = device.Decon(shape, 1)
d.set_psf(0, psf)
g = torch.Generator(device="cuda"); g.manual_seed(1)
d.set_image(0, torch.rand(shape, device="cuda", generator=g) * 100 + 10)
d.run(5); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); d.run(30); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 30)
d.run(10)
chk = float(torch.as_tensor(d.result()).double().sum())
print(os.environ.get("AB_TAG", ""), "ms/iter %.4f" % best, "checksum %.6f" % chk, "kernels", [round(float(x) * 1e3, 1) for x in d.time_kernels(5)])
