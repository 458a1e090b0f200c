"""Multi-process host logic on CPU: world_size 2 over gloo.  Time points of spimFusionBatch are
independent units; ranks take them round robin and only timings are reduced."""
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
from microimagelib_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard.shard_time_points(3, 40, 2, rank, world)
# gather everybody's share on every rank
got = [None] * world
dist.all_gather_object(got, mine)
t = shard.max_over_ranks(1.0 + rank)           # "time" = slowest rank
n = shard.sum_over_ranks(len(mine))            # units processed by the whole job
dist.barrier()
if rank == 0:
    print(json.dumps({"shares": got, "tmax": t, "units": n}))
dist.destroy_process_group()
""" % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_round_robin_rule():
    from microimagelib_b200 import shard
    allp = shard.time_points(0, 63, 1)
    parts = [shard.shard_time_points(0, 63, 1, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == allp and all(len(p) == 8 for p in parts)
    assert shard.shard_time_points(5, 5, 1, 0, 2) == [5] and shard.shard_time_points(5, 5, 1, 1, 2) == []
    with pytest.raises(ValueError):
        shard.shard_time_points(0, 3, 1, 2, 2)


def test_two_ranks_over_gloo(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    a, b = out["shares"]
    assert sorted(a + b) == list(range(3, 41, 2)) and not set(a) & set(b)
    assert abs(len(a) - len(b)) <= 1
    assert out["tmax"] == 2.0 and out["units"] == len(a) + len(b)
