"""Host-side logic of the slab-decomposed distributed FFT (microimagelib_b200/dist_decon.py) on CPU:
the plane split and the two all-to-all re-layouts, world_size 2 over gloo."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from microimagelib_b200.dist_decon import SlabLayout
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
X, Y, Z = 16, 8, 6
L = SlabLayout((X, Y, Z), rank, world)
kx = torch.arange(L.nplanes).view(-1, 1, 1, 1).float()
y = (torch.arange(L.ny) + L.y0).view(1, -1, 1, 1).float()
z = torch.arange(Z).view(1, 1, -1, 1).float()
c = torch.tensor([0.0, 0.5]).view(1, 1, 1, 2)
slab = (kx * 10000 + y * 100 + z + c).contiguous()                     # value encodes (kx, y, z, re/im)
planes = torch.zeros((L.np, Y, Z, 2))
scratch = torch.zeros(world * max(L.counts) * L.row)
L.to_planes(slab, planes, scratch)
kxp = (torch.arange(L.np) + L.p0).view(-1, 1, 1, 1).float()
yy = torch.arange(Y).view(1, -1, 1, 1).float()
want = kxp * 10000 + yy * 100 + z + c
ok1 = bool(torch.equal(planes, want))
back = torch.zeros_like(slab)
L.to_slabs(planes, back, scratch)
ok2 = bool(torch.equal(back, slab))
res = [None] * world
dist.all_gather_object(res, dict(rank=rank, np=L.np, p0=L.p0, ok1=ok1, ok2=ok2))
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
""" % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_plane_counts():
    from microimagelib_b200.dist_decon import plane_counts
    assert plane_counts(257, 8) == [33] + [32] * 7
    assert plane_counts(129, 2) == [65, 64]
    assert sum(plane_counts(513, 4)) == 513


def test_single_rank_layout_roundtrip():
    import torch
    from microimagelib_b200.dist_decon import SlabLayout
    L = SlabLayout((8, 4, 6), 0, 1)
    slab = torch.randn((L.nplanes, L.ny, L.Z, 2))
    planes = torch.zeros((L.np, L.Y, L.Z, 2))
    scratch = torch.zeros(L.np * L.row)
    L.to_planes(slab, planes, scratch)
    assert torch.equal(planes, slab)
    back = torch.zeros_like(slab)
    L.to_slabs(planes, back, scratch)
    assert torch.equal(back, slab)


def test_two_rank_relayout_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=280)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    assert [o["np"] for o in out] == [5, 4] and [o["p0"] for o in out] == [0, 5]
    assert all(o["ok1"] and o["ok2"] for o in out)
