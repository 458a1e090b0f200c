"""CPU restatement of reg3d's affine path (warp + ZNCC cost + Powell schedule).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Numerics live in ``reg_oracle.c`` (built to ``oracle/liboracle_reg.so`` by ``oracle/Makefile``);
the optimiser is the reference's own ``src/api_powell.c`` compiled unchanged into
``oracle/_ref/libpowell_ref.so``.  Volumes are numpy C-order ``(slices, H, W)`` float32, i.e. the
reference's ``x + y*sx + z*sx*sy`` layout with sx = W.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_F = C.POINTER(C.c_float)
_LL = C.c_longlong

NDIM = 12


def _load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
    return C.CDLL(path)


_lib = None
_pow = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load(os.path.join(_HERE, "liboracle_reg.so"))
        _lib.orc_affine_warp.argtypes = [_F, _F, _LL, _LL, _LL, _LL, _LL, _LL, _F]
        _lib.orc_zncc_sums.argtypes = [_F, _F, _LL, _LL, _LL, _LL, _LL, _LL, _F,
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.orc_cost_from_sums.argtypes = [C.c_double, C.c_double, C.c_float]
        _lib.orc_cost_from_sums.restype = C.c_float
        _lib.orc_sum.argtypes = [_F, _LL]
        _lib.orc_sum.restype = C.c_double
        _lib.orc_demean.argtypes = [_F, _F, _LL]
        _lib.orc_demean.restype = C.c_float
        _lib.orc_p2matrix.argtypes = [_F, _F]
        _lib.orc_matrix2p.argtypes = [_F, _F]
        _lib.orc_matrixmultiply.argtypes = [_F, _F, _F]
        _lib.orc_dof9tomatrix.argtypes = [_F, _F, C.c_int]
    return _lib


COSTFN = C.CFUNCTYPE(C.c_float, _F)


def powell_ref():
    """The reference optimiser, src/api_powell.c:305, compiled unchanged."""
    global _pow
    if _pow is None:
        _pow = _load(os.path.join(_HERE, "_ref", "libpowell_ref.so"))
        _pow.powell.argtypes = [_F, C.POINTER(_F), C.c_int, C.c_float, C.POINTER(C.c_int), _F,
                                COSTFN, C.POINTER(C.c_int), C.c_int]
        _pow.powell.restype = None
    return _pow


def _fp(a):
    return a.ctypes.data_as(_F)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ------------------------------------------------------------------ kernels
def affine_warp(src, aff, out_shape=None):
    """a17 affineTransform (include/cukernel.cuh:500-524); aff maps output voxel -> source voxel."""
    src = _f32(src)
    aff = _f32(aff).reshape(12)
    out_shape = tuple(out_shape or src.shape)
    out = np.empty(out_shape, dtype=np.float32)
    sz, sy, sx = out_shape
    sz2, sy2, sx2 = src.shape
    lib().orc_affine_warp(_fp(out), _fp(src), sx, sy, sz, sx2, sy2, sz2, _fp(aff))
    return out


def zncc_sums(target, src, aff):
    """a14 corrkernel double sums (sum s*s, sum s*t)."""
    target = _f32(target)
    src = _f32(src)
    aff = _f32(aff).reshape(12)
    sz, sy, sx = target.shape
    sz2, sy2, sx2 = src.shape
    ss, st = C.c_double(), C.c_double()
    lib().orc_zncc_sums(_fp(target), _fp(src), sx, sy, sz, sx2, sy2, sz2, _fp(aff),
                        C.byref(ss), C.byref(st))
    return ss.value, st.value


def cost_from_sums(ss, st, sd_t):
    return float(lib().orc_cost_from_sums(ss, st, np.float32(sd_t)))


def demean(vol):
    """(vol - mean, sqrt(sum sq)) with the reference's float/double mix, src/api_subfunc.cu:2838-2868."""
    vol = _f32(vol)
    out = np.empty_like(vol)
    sd = lib().orc_demean(_fp(out), _fp(vol), vol.size)
    return out, np.float32(sd)


def zncc_cost(target_dm, sd_t, src_dm, aff):
    """costfunc value (= -ZNCC) for pre-demeaned volumes, src/api_subfunc.cu:954-988, 2377-2388."""
    ss, st = zncc_sums(target_dm, src_dm, aff)
    return cost_from_sums(ss, st, sd_t)


# ------------------------------------------------------------------ parameter <-> matrix
def p2matrix(x):
    x = _f32(x)
    m = np.zeros(12, np.float32)
    lib().orc_p2matrix(_fp(m), _fp(x))
    return m


def matrix2p(m):
    m = _f32(m)
    x = np.zeros(13, np.float32)
    lib().orc_matrix2p(_fp(m), _fp(x))
    return x


def matrixmultiply(m1, m2):
    m1, m2 = _f32(m1).reshape(12), _f32(m2).reshape(12)
    m = np.zeros(12, np.float32)
    lib().orc_matrixmultiply(_fp(m), _fp(m1), _fp(m2))
    return m


def dof9tomatrix(p_dof, dof_num):
    p_dof = _f32(p_dof)
    assert p_dof.size >= 10
    m = np.zeros(12, np.float32)
    lib().orc_dof9tomatrix(_fp(m), _fp(p_dof), int(dof_num))
    return m


def checkmatrix(m, sx, sy, sz):
    """src/api_reg.cpp:247-262."""
    m = _f32(m).reshape(12)
    ok = True
    lo, up = np.float32(0.5), np.float32(1.4)
    if m[0] < lo or m[0] > up or m[5] < lo or m[5] > up or m[10] < lo or m[10] > up:
        ok = False
    s = np.float32(np.float32(m[0] + m[5]) + m[10])
    if s < np.float32(2) or s > np.float32(4):
        ok = False
    r = np.float32(0.8)
    if abs(m[3]) > r * np.float32(sx) or abs(m[7]) > r * np.float32(sy) or abs(m[11]) > r * np.float32(sz):
        ok = False
    return ok


# ------------------------------------------------------------------ Powell driver (reference optimiser)
def run_powell_ref(p, xi, n, ftol, func, it_limit, cit=None):
    """Call the reference's powell() (src/api_powell.c:305) on the 1-indexed vector p[0..n] and the
    direction set xi (list of rows, leading n x n block used; updated in place).
    `func(x)` gets a copy of the 1-indexed trial vector.  `cit` is the shared evaluation counter
    (*totalIt): the reference's costfunc increments it itself, so `func` must do `cit.value += 1`.
    Returns (iter, fret)."""
    lp = powell_ref()
    rows = (_F * (n + 1))()
    store = []
    for i in range(1, n + 1):
        row = (C.c_float * (n + 1))(*([0.0] + [float(xi[i - 1][j]) for j in range(n)]))
        store.append(row)
        rows[i] = C.cast(row, _F)
    cit = cit if cit is not None else C.c_int(0)
    it = C.c_int(0)
    fr = C.c_float(0)

    def _cb(ptr):
        x = np.ctypeslib.as_array(ptr, shape=(n + 1,)).copy()
        return float(func(x))

    cb = COSTFN(_cb)
    lp.powell(_fp(p), rows, n, C.c_float(ftol), C.byref(it), C.byref(fr), cb, C.byref(cit), int(it_limit))
    for i in range(n):
        for j in range(n):
            xi[i][j] = store[i][j + 1]
    return it.value, np.float32(fr.value)


def reg3d_affine(target, source, aff_method, flag_tmx=False, itmx=None, ftol=1e-4, it_limit=3000,
                 trace=None):
    """reg3d_affine1, src/api_subfunc.cu:2733-2994, driven by the reference's powell().
    Returns dict(reg=..., tmx=..., records=...).  `trace`, if a list, receives every evaluated
    (matrix, cost) pair in order."""
    target = _f32(target)
    source = _f32(source)
    itmx = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32) if itmx is None else _f32(itmx).reshape(12).copy()
    records = np.zeros(11, np.float32)
    if aff_method == 0:
        if flag_tmx:
            reg = affine_warp(source, itmx, target.shape)
        else:
            reg = source.copy()
            itmx = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
        return dict(reg=reg, tmx=itmx, records=records)

    aff_initial = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    pre = source
    if flag_tmx:
        if aff_method == 5:
            aff_initial = itmx.copy()
        else:
            pre = affine_warp(source, itmx, target.shape)
    src_dm, _ = demean(pre)
    tgt_dm, sd_t = demean(target)
    if sd_t == 0:
        raise ValueError("SD of image 1 is zero")

    state = dict(aff=np.zeros(12, np.float32), n_eval=0, dof9=False, dof=12)

    def costfunc(x):
        if state["dof9"]:
            state["aff"] = dof9tomatrix(np.concatenate([x, np.zeros(max(0, 10 - x.size), np.float32)]), state["dof"])
        else:
            state["aff"] = p2matrix(x)
        c = zncc_cost(tgt_dm, sd_t, src_dm, state["aff"])
        state["n_eval"] += 1
        if trace is not None:
            trace.append((state["aff"].copy(), np.float32(c)))
        return c

    p = matrix2p(aff_initial)
    records[1] = -np.float32(costfunc(p))
    state["n_eval"] = 0
    xi12 = [[1.0 if i == j else 0.0 for j in range(12)] for i in range(12)]
    xi9 = [[1.0 if i == j else 0.0 for j in range(9)] for i in range(9)]
    p9 = np.array([0, 0, 0, 0, 0, 0, 0, 1, 1, 1], np.float32)
    fret = np.float32(0)

    cit = C.c_int(0)

    def run_shared(pvec, xi, n, tol):
        nonlocal fret
        cit.value = state["n_eval"]
        _, fret = run_powell_ref(pvec, xi, n, tol, counted, it_limit, cit)

    def counted(x):
        v = costfunc(x)
        cit.value = state["n_eval"]
        return v

    if aff_method in (1, 2, 3, 4):
        state["dof9"], state["dof"] = True, {1: 3, 2: 6, 3: 7, 4: 9}[aff_method]
        run_shared(p9, xi9, state["dof"], ftol)
    elif aff_method == 5:
        state["dof9"], state["dof"] = False, 12
        run_shared(p, xi12, 12, ftol)
    elif aff_method == 6:
        state["dof9"], state["dof"] = True, 6
        run_shared(p9, xi9, 6, 0.01)
        records[2] = -fret
        state["dof9"], state["dof"] = False, 12
        p = matrix2p(state["aff"])
        run_shared(p, xi12, 12, ftol)
    elif aff_method == 7:
        state["dof9"] = True
        for d, tol in ((3, 0.01), (6, 0.01), (9, 0.005)):
            state["dof"] = d
            run_shared(p9, xi9, d, tol)
        records[2] = -fret
        state["dof9"], state["dof"] = False, 12
        p = matrix2p(state["aff"])
        run_shared(p, xi12, 12, ftol)
    else:
        raise ValueError("bad affMethod")

    aff = state["aff"].copy()
    if flag_tmx and aff_method != 5:
        aff = matrixmultiply(itmx, aff)
    records[3] = -fret
    records[5] = state["n_eval"]
    reg = affine_warp(source, aff, target.shape)
    return dict(reg=reg, tmx=aff, records=records)
