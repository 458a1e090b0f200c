"""The product's Powell optimiser (csrc/powell.cpp, host only) against the reference's own
src/api_powell.c compiled unchanged into oracle/_ref/libpowell_ref.so: every evaluated point and
every returned number must be bit-identical, because reg3d returns the matrix of the LAST
evaluated point."""
import ctypes as C
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "powell_trajectories.json")


def _funcs():
    f32 = np.float32

    def quad(x):       # separable bowl, minimum -0.8 at (1, -2, 0.5, ...)
        c = np.array([1, -2, 0.5, 3, -1, 0.25, 2, -0.5, 1.5, -3, 0.75, -1.25], f32)[: len(x) - 1]
        return f32(np.sum((x[1:] - c) ** 2, dtype=f32) * f32(0.01) - f32(0.8))

    def rosen(x):      # curved valley
        s = f32(0)
        for i in range(1, len(x) - 1):
            s += f32(100) * (x[i + 1] - x[i] * x[i]) ** 2 + (f32(1) - x[i]) ** 2
        return f32(s * f32(1e-3) - f32(0.9))

    def zncc_like(x):  # smooth bump in [-1, 0], like -ZNCC around an optimum
        c = np.array([0.3, -0.2, 0.1, 1.0, 0.02, -0.01, 0.0, 1.01, 0.0, 0.03, 0.0, 0.99], f32)[: len(x) - 1]
        d = x[1:] - c
        return f32(-np.exp(-np.sum(d * d, dtype=f32) * f32(0.5)))

    return {"quad": quad, "rosen": rosen, "zncc_like": zncc_like}


CASES = [("quad", 3, 1e-4, 3000), ("quad", 12, 1e-4, 3000), ("rosen", 4, 1e-4, 3000), ("zncc_like", 6, 0.01, 3000),
         ("zncc_like", 12, 1e-4, 3000), ("zncc_like", 12, 1e-4, 40), ("rosen", 9, 0.005, 500)]


def _start(name, n):
    if name == "zncc_like":
        p = np.zeros(n + 1, np.float32)
        if n == 12:
            p[[4, 8, 12]] = 1
        return p
    return np.zeros(n + 1, np.float32)


def _run(impl, name, n, ftol, limit):
    f = _funcs()[name]
    trace = []
    cnt = C.c_int(0)

    def cost(x):
        v = f(x)
        trace.append((x[1:].copy(), np.float32(v)))
        cnt.value += 1
        return v

    p = _start(name, n)
    xi = np.eye(n, dtype=np.float32)
    if impl == "ref":
        from oracle import reg_oracle as ro
        xil = [list(r) for r in xi]
        it, fret = ro.run_powell_ref(p, xil, n, ftol, cost, limit, cnt)
        xi = np.array(xil, np.float32)
    else:
        from microimagelib_b200 import device
        it, fret, p = device.powell(p, xi, n, ftol, cost, limit, cnt)
    return dict(iter=it, fret=np.float32(fret), p=p.copy(), xi=xi, trace=trace)


@pytest.mark.parametrize("name,n,ftol,limit", CASES)
def test_matches_reference_bit_for_bit(name, n, ftol, limit):
    if not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libpowell_ref.so")):
        pytest.skip("oracle/_ref not built (reference tree absent and no prebuilt copy)")
    a = _run("ref", name, n, ftol, limit)
    b = _run("milb", name, n, ftol, limit)
    assert len(a["trace"]) == len(b["trace"])
    for (xa, va), (xb, vb) in zip(a["trace"], b["trace"]):
        assert np.array_equal(xa.view(np.uint32), xb.view(np.uint32))
        assert va == vb
    assert a["iter"] == b["iter"]
    assert a["fret"] == b["fret"]
    assert np.array_equal(a["p"][1:].view(np.uint32), b["p"][1:].view(np.uint32))
    assert np.array_equal(a["xi"].view(np.uint32), b["xi"].view(np.uint32))


@pytest.mark.parametrize("name,n,ftol,limit", CASES)
def test_matches_committed_golden_trajectories(name, n, ftol, limit):
    """Golden vectors written by tests/golden/make_powell_golden.py from the reference optimiser."""
    with open(GOLD) as fh:
        gold = json.load(fh)
    g = gold[f"{name}-{n}-{ftol}-{limit}"]
    b = _run("milb", name, n, ftol, limit)
    assert len(b["trace"]) == g["n_eval"]
    assert int(np.float32(b["fret"]).view(np.uint32)) == g["fret_bits"]
    assert [int(v) for v in b["p"][1:].view(np.uint32)] == g["p_bits"]
    last = b["trace"][-1][0]
    assert [int(v) for v in last.view(np.uint32)] == g["last_x_bits"]
    import hashlib
    h = hashlib.sha256()
    for x, v in b["trace"]:
        h.update(x.tobytes()); h.update(np.float32(v).tobytes())
    assert h.hexdigest() == g["trace_sha256"]
