"""The slab-decomposed distributed deconvolution against the single-GPU path: same kernels, same
arithmetic order per element, so the results must be bit-identical.  The 1-rank case runs in the
normal GPU suite; the 2-rank NCCL case needs two GPUs (gpurun --gpus 2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from microimagelib_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(shape, dual):
    psf_a = synth.gaussian_psf((17, 17, 17), (3, 2, 2))
    psf_b = synth.gaussian_psf((17, 17, 17), (2, 2, 3))
    a = synth.bead_image(shape, psf_a, density=1 / 4096.0)
    b = synth.bead_image(shape, psf_b, density=1 / 4096.0, noise_seed=5)
    return (a, b, psf_a, psf_b) if dual else (a, None, psf_a, None)


def _single_gpu(shape, dual, iters):
    from microimagelib_b200 import device
    a, b, pa, pb = _inputs(shape, dual)
    d = device.Decon(shape, 2 if dual else 1)
    d.set_psf(0, pa)
    d.set_image(0, a)
    if dual:
        d.set_psf(1, pb)
        d.set_image(1, b)
    d.run(iters)
    out = d.result().copy()
    d.close()
    return out


@pytest.mark.parametrize("shape", [(64, 128, 128), (64, 1024, 128)])   # the second: Y = 1024, k_ypassW and its peer-store variant
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("dual", [False, True])
def test_one_rank_distributed_path_equals_single_gpu_path(dual, fused, shape):
    import torch
    from microimagelib_b200.dist_decon import DistDecon
    a, b, pa, pb = _inputs(shape, dual)
    dd = DistDecon(shape, 2 if dual else 1, fused=fused)
    assert dd.fused == fused
    dd.set_psf(0, pa)
    dd.set_image(0, a)
    if dual:
        dd.set_psf(1, pb)
        dd.set_image(1, b)
    got = dd.run(4).cpu().numpy()
    torch.cuda.synchronize()
    assert np.array_equal(got, _single_gpu(shape, dual, 4))
    dd.close()


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%r, "tests"))
import test_gpu_dist_decon as T
from microimagelib_b200.dist_decon import DistDecon
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dual, iters = True, 3
for shape, fused in (((64, 128, 128), False), ((64, 128, 128), True), ((64, 1024, 128), True)):
    a, b, pa, pb = T._inputs(shape, dual)
    ref = T._single_gpu(shape, dual, iters) if rank == 0 else None
    dd = DistDecon(shape, 2, fused=fused)
    assert dd.fused == fused
    L = dd.L
    dd.set_psf(0, pa); dd.set_psf(1, pb)
    dd.set_image(0, a[:, L.y0:L.y0 + L.ny, :]); dd.set_image(1, b[:, L.y0:L.y0 + L.ny, :])
    for rep in range(2):                         # twice: the second run reuses the mapped buffers
        E = dd.run(iters)
    parts = [torch.empty_like(E) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, E)
    if rank == 0:
        got = torch.cat(parts, dim=1).cpu().numpy()
        print("DIST_EQUAL", "fused" if fused else "nccl", bool(np.array_equal(got, ref)), float(np.abs(got - ref).max()))
    dd.close()
dist.destroy_process_group()
""" % (ROOT, ROOT)


def test_two_ranks_nccl_and_fused_exchange(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("DIST_EQUAL")]
    assert len(lines) == 3, r.stdout[-1500:] + r.stderr[-3000:]
    for line in lines:                            # the NCCL all-to-all path, the fused peer-store path, and the latter at Y = 1024
        assert line.split()[2] == "True", line
