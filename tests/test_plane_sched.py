"""Host replay of the fused plane stage's pipeline (csrc/plane_sched.h, compiled with g++ into the
emulation library): the CTA roles cover every (phase, plane, tile) exactly once, and a discrete
replay of the kernel's rules -- a CTA takes its role's tiles in order, a tile starts only when its
dependency counter is complete, completion is signalled one tile late -- always runs to the end
(no deadlock) and never lets two planes share a ring slot."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emul", "libfft_emul.so")


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "emul", "fft_emul.cpp")
    dep = os.path.join(HERE, "..", "microimagelib_b200", "csrc", "plane_sched.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", SO, src], check=True)
    return C.CDLL(SO)


CASES = [  # planes, tiles per plane, ring, grid, nA, nB
    (129, 32, 16, 148, 42, 68), (257, 32, 16, 148, 42, 68), (65, 8, 4, 148, 40, 70), (33, 4, 2, 296, 90, 130),
    (513, 128, 8, 148, 42, 68), (5, 32, 16, 148, 50, 50), (1, 2, 2, 3, 1, 1), (129, 32, 2, 7, 2, 3),
]


@pytest.mark.parametrize("planes,tpp,ring,grid,nA,nB", CASES)
def test_roles_cover_everything_and_the_pipeline_drains(lib, planes, tpp, ring, grid, nA, nB):
    roles = np.zeros((grid, 3), np.int32)
    lib.emul_plane_roles(grid, nA, nB, roles.ctypes.data_as(C.POINTER(C.c_int)))
    ntiles = planes * tpp
    seen = set()
    queues = []                      # per CTA: (phase, [tile indices j])
    for b in range(grid):
        ph, rank, nrole = roles[b].tolist()
        assert 0 <= ph < 3 and 0 <= rank < nrole
        js = list(range(rank, ntiles, nrole))
        for j in js:
            assert (ph, j) not in seen
            seen.add((ph, j))
        queues.append((ph, js))
    assert len(seen) == 3 * ntiles

    def dep(ph, plane):
        dp = C.c_int(0)
        k = lib.emul_plane_dependency(ph, plane, ring, C.byref(dp))
        return k, dp.value

    done = {0: [0] * planes, 1: [0] * planes}       # signalled tiles per plane of phase A / B
    pos = [0] * grid                                 # next tile of each CTA
    pending = [None] * grid                          # (phase, plane) finished but not yet signalled
    slot_owner = {}                                  # ring slot -> plane currently written / not yet consumed
    remaining = 3 * ntiles
    rounds = 0
    while remaining:
        progressed = False
        for b in range(grid):
            ph, js = queues[b]
            if pos[b] >= len(js):
                if pending[b]:
                    done[pending[b][0]][pending[b][1]] += 1
                    pending[b] = None
                    progressed = True
                continue
            plane = js[pos[b]] // tpp
            kind, dp = dep(ph, plane)
            ok = kind == 0 or done[kind - 1][dp] == tpp
            if not ok:
                if pending[b]:                       # a waiting CTA sends its pending signal first
                    done[pending[b][0]][pending[b][1]] += 1
                    pending[b] = None
                    progressed = True
                continue
            if ph == 0:                              # phase A writes ring slot plane % ring
                s = plane % ring
                if s in slot_owner and slot_owner[s] != plane:
                    assert done[1][slot_owner[s]] == tpp, "ring slot rewritten before it was consumed"
                slot_owner[s] = plane
            if pending[b]:                           # the previous tile is signalled during this one
                done[pending[b][0]][pending[b][1]] += 1
            pending[b] = (ph, plane) if ph < 2 else None
            pos[b] += 1
            remaining -= 1
            progressed = True
        rounds += 1
        assert progressed, "pipeline deadlocked"
        assert rounds < 20 * (ntiles + grid)
    for b in range(grid):
        if pending[b]:
            done[pending[b][0]][pending[b][1]] += 1
    assert all(v == tpp for v in done[0]) and all(v == tpp for v in done[1])
