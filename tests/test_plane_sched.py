"""Host check of the fused plane stage's work list (csrc/plane_sched.h, compiled with g++ into the
emulation library): every (phase, plane, tile) appears exactly once, every ticket's producers come
earlier in the list (so the static round-robin deal cannot deadlock), and a ring slot is only
rewritten after its previous plane has been consumed."""
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


@pytest.mark.parametrize("planes,group,tpp,ring", [(129, 4, 32, 16), (257, 4, 32, 16), (65, 16, 8, 64), (33, 64, 2, 128), (513, 1, 128, 2),
                                                    (129, 5, 32, 10), (3, 4, 32, 16), (1, 1, 1, 2)])
def test_schedule_is_complete_and_ordered(lib, planes, group, tpp, ring):
    cap = (planes // group + 4) * 3 * group * tpp
    out = np.zeros((cap, 6), np.int32)
    total = lib.emul_plane_tickets(planes, group, tpp, ring, out.ctypes.data_as(C.POINTER(C.c_int)), cap)
    assert 0 < total <= cap
    out = out[:total]
    seen = {}
    last_of = {}           # (phase, plane) -> ticket of its last tile
    for t, (ok, ph, pl, ti, kind, dp) in enumerate(out.tolist()):
        if not ok:
            continue
        assert 0 <= ph < 3 and 0 <= pl < planes and 0 <= ti < tpp
        assert (ph, pl, ti) not in seen
        seen[(ph, pl, ti)] = t
        last_of[(ph, pl)] = t
    assert len(seen) == 3 * planes * tpp
    for (ph, pl, ti), t in seen.items():
        ok, _, _, _, kind, dp = out[t].tolist()
        if ph == 0:
            assert (kind, dp) == ((2, pl - ring) if pl >= ring else (0, pl - ring))
        else:
            assert (kind, dp) == (ph, pl)
        if kind:
            prod = (kind - 1, dp)                    # phase A (0) or phase B (1) of plane dp
            assert last_of[prod] < t, "a ticket must come after every ticket it waits for"
    # ring: between A(p) and the end of B(p) no other plane may be written into slot p mod ring
    for p in range(planes):
        a0 = min(seen[(0, p, i)] for i in range(tpp))
        b1 = last_of[(1, p)]
        for q in range(p % ring, planes, ring):
            if q == p:
                continue
            qa = [seen[(0, q, i)] for i in range(tpp)]
            assert max(qa) < a0 or min(qa) > a0, "two planes interleave in one ring slot"
            if q > p:
                # q's phase A waits for B(q - ring) ... B(p) transitively: its dependency chain reaches p
                assert (q - p) % ring == 0
