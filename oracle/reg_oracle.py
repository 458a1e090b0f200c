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
        _lib.orc_corr2d_sums.argtypes = [_F, _F, _LL, _LL, _LL, _LL, _F, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.orc_corr2d_costs.argtypes = [_F, _F, _LL, _LL, _LL, _LL, _F, C.c_int, C.c_float, _F]
        _lib.orc_affine2d.argtypes = [_F, _F, _LL, _LL, _LL, _LL, _F]
        _lib.orc_tex2d_samples.argtypes = [_F, _F, _LL, _LL, _F, _LL]
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


# ------------------------------------------------------------------ pre-alignment (regChoice 1 / 3 / 4, reg2d)
def tex2d_samples(img, coords):
    """tex2D (linear filter, un-normalised coordinates) of a (H, W) image at (n, 2) texture coordinates."""
    img = _f32(img)
    coords = _f32(coords).reshape(-1, 2)
    out = np.empty(coords.shape[0], np.float32)
    lib().orc_tex2d_samples(_fp(out), _fp(img), img.shape[1], img.shape[0], _fp(coords), coords.shape[0])
    return out


def corr2d_costs(tgt_dm, sd_t, src_dm, affs):
    """costfunc2D (= -ZNCC, +2 when the warped image is empty) for K 2x3 matrices, src/api_subfunc.cu:1014-1036, 1815-1821."""
    tgt_dm = _f32(tgt_dm)
    src_dm = _f32(src_dm)
    affs = _f32(affs).reshape(-1, 6)
    out = np.empty(affs.shape[0], np.float32)
    lib().orc_corr2d_costs(_fp(tgt_dm), _fp(src_dm), tgt_dm.shape[1], tgt_dm.shape[0], src_dm.shape[1], src_dm.shape[0],
                           _fp(affs), affs.shape[0], C.c_float(float(sd_t)), _fp(out))
    return out


def affine2d(src, aff, out_shape):
    """affineTransform2D, include/cukernel.cuh:558-573"""
    src = _f32(src)
    out = np.empty(tuple(out_shape), np.float32)
    lib().orc_affine2d(_fp(out), _fp(src), out.shape[1], out.shape[0], src.shape[1], src.shape[0], _fp(_f32(aff).reshape(6)))
    return out


def _init_aff2d(img1, img2, flag_tmx, itmx):
    if flag_tmx:
        return _f32(itmx).reshape(6).copy()
    (sy, sx), (sy2, sx2) = img1.shape, img2.shape
    return np.array([1, 0, int((sx2 - sx) / 2), 0, 1, int((sy2 - sy) / 2)], np.float32)


def reg2d_shiftalign(img1, img2, flag_tmx=False, itmx=None, search_y=True, shift_region=0.3, total_step=30.0):
    """reg2d_shiftalign1 (search_y) / reg2d_shiftalignX1 (x only), src/api_subfunc.cu:1860-2117.
    Images are (H, W).  Returns dict(tmx, initial, best, reg)."""
    img1, img2 = _f32(img1), _f32(img2)
    aff = _init_aff2d(img1, img2, flag_tmx, itmx)
    tgt_dm, sd_t = demean(img1)
    if sd_t == 0:
        raise ValueError("SD of image 1 is zero")
    src_dm, _ = demean(img2)
    f32 = np.float32
    steps = int(total_step)
    off_x, off_y = aff[2], aff[5]
    step_x = f32(f32(f32(img2.shape[1]) * f32(shift_region)) / f32(total_step))
    step_y = f32(f32(f32(img2.shape[0]) * f32(shift_region)) / f32(total_step))
    cands = [(off_x, off_y)]
    for i in range(-steps, steps):
        px = f32(off_x + f32(step_x * f32(i)))
        if search_y:
            for j in range(-steps, steps):
                cands.append((px, f32(off_y + f32(step_y * f32(j)))))
        else:
            cands.append((px, off_y))
    mats = np.array([[aff[0], aff[1], x, aff[3], aff[4], y] for x, y in cands], np.float32)
    costs = corr2d_costs(tgt_dm, sd_t, src_dm, mats)
    best, sx_, sy_ = f32(0), f32(0), (f32(0) if search_y else off_y)
    for k in range(1, len(cands)):
        v = f32(-costs[k])
        if v > best:
            best, sx_, sy_ = v, mats[k, 2], mats[k, 5]
    aff[2], aff[5] = sx_, sy_
    fin = corr2d_costs(tgt_dm, sd_t, src_dm, aff[None])[0]
    return dict(tmx=aff, initial=f32(-costs[0]), best=f32(-fin), reg=affine2d(src_dm, aff, img1.shape))


def reg2d_affine(img1, img2, aff_method=1, flag_tmx=False, itmx=None, ftol=1e-4, it_limit=3000):
    """reg2d_affine1, src/api_subfunc.cu:2229-2336, driven by the reference's powell()."""
    img1, img2 = _f32(img1), _f32(img2)
    aff = _init_aff2d(img1, img2, flag_tmx, itmx)
    tgt_dm, sd_t = demean(img1)
    if sd_t == 0:
        raise ValueError("SD of image 1 is zero")
    src_dm, _ = demean(img2)
    state = dict(last=aff.copy())
    cit = C.c_int(0)

    def cost(x):
        state["last"] = _f32(x[1:7]).copy()
        cit.value += 1
        return corr2d_costs(tgt_dm, sd_t, src_dm, state["last"][None])[0]

    p = np.concatenate([[0], aff]).astype(np.float32)
    first = np.float32(-cost(p))
    fret = np.float32(0)
    tmx = _f32(itmx).reshape(6).copy() if itmx is not None else aff.copy()
    if aff_method > 0:
        xi = [[1.0 if i == j else 0.0 for j in range(6)] for i in range(6)]
        _, fret = run_powell_ref(p, xi, 6, ftol, cost, it_limit, cit)
        tmx = state["last"].copy()
    return dict(tmx=tmx, initial=first, best=np.float32(-fret), n_eval=cit.value, reg=affine2d(img2, state["last"], img1.shape))


def max_projection(vol, direction):
    """maxprojectionkernel, include/cukernel.cuh:396-418 (the running maximum starts at 0).
    vol is (S, H, W); direction 1 -> (H, W) image, 2 -> (W, S) image [x rows, z fastest], 3 -> (S, H)."""
    vol = _f32(vol)
    z = np.float32(0)
    if direction == 1:
        return np.maximum(vol.max(axis=0), z)
    if direction == 2:
        return np.ascontiguousarray(np.maximum(vol.max(axis=1), z).T)     # (S, W) -> rows x, columns z
    return np.ascontiguousarray(np.maximum(vol.max(axis=2), z))            # out[i=y + j*sy], j = z


def prealign_mip(target, source):
    """reg3d regChoice 4 pre-alignment, src/api_reg.cpp:466-501: returns the 12-element translation matrix."""
    r1 = reg2d_shiftalign(max_projection(target, 1), max_projection(source, 1), False, None, True, 0.3, 30.0)
    tmx1 = r1["tmx"]
    tmx2 = np.array([1, 0, 0, 0, 1, tmx1[2]], np.float32)
    r2 = reg2d_shiftalign(max_projection(target, 2), max_projection(source, 2), True, tmx2, False, 0.3, 30.0)
    m = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    m[3], m[7], m[11] = tmx1[2], tmx1[5], r2["tmx"][2]
    return m


def _zncc1(a, b):
    """zncc1, src/api_subfunc.cu:2409-2432 (float mean shift and products, double sums)."""
    a, b = _f32(a).ravel(), _f32(b).ravel()
    n = np.float32(a.size)
    a1 = a + (-np.float32(a.sum(dtype=np.float64)) / n)
    b1 = b + (-np.float32(b.sum(dtype=np.float64)) / n)
    st = (a1 * b1).sum(dtype=np.float64)
    tt = (a1 * a1).sum(dtype=np.float64)
    ss = (b1 * b1).sum(dtype=np.float64)
    d = np.float32(np.sqrt(tt * ss))
    return np.float32(st / np.float64(d)) if d != 0 else np.float32(-2.0)


def phasor(img1, img2):
    """reg3d_phasor1 / reg2d_phasor1 (src/api_subfunc.cu:2466-2590, 2128-2227): integer shift (x, y, z)
    of img2 against img1 by phase correlation.  Volumes (S, H, W); a (H, W) image is treated as S = 1."""
    img1, img2 = _f32(img1), _f32(img2)
    if img1.ndim == 2:
        img1, img2 = img1[None], img2[None]
    sz, sy, sx = img1.shape
    f1 = np.fft.rfftn(img1.astype(np.float64))
    f2 = np.fft.rfftn(img2.astype(np.float64))
    c = np.conj(f1) * f2
    e = np.abs(c)
    c = np.where(e != 0, c / np.where(e != 0, e, 1), 0)
    ph = np.fft.irfftn(c, s=img1.shape, axes=(0, 1, 2)).astype(np.float32)
    ph = np.roll(ph, (sz // 2, sy // 2, sx // 2), axis=(0, 1, 2))
    # max3Dgpu (:437-470): first z of each column, then first column in x-outer / y-inner order
    colmax = ph.max(axis=0)
    colz = ph.argmax(axis=0)
    order = colmax.T                                     # [x][y]
    flat = int(np.argmax(order))                          # first maximum in x-major order
    cx, cy = divmod(flat, sy)
    cz = 0 if (cx == 0 and cy == 0) else int(colz[cy, cx])
    shift = [cx - sx // 2, cy - sy // 2, cz - sz // 2]
    dims = [sx, sy, sz]
    ab = [abs(v) for v in shift]
    if any(ab[d] > dims[d] // 4 for d in range(3)):
        img_t = np.roll(img2, (-shift[2], -shift[1], -shift[0]), axis=(0, 1, 2))
        crop = [(dims[d] - ab[d], ab[d]) for d in range(3)]
        org = [((0, dims[d] - ab[d]) if shift[d] > 0 else (ab[d], 0)) for d in range(3)]
        best, ind = np.float32(-3), (0, 0, 0)
        for i in range(2):
            if not crop[0][i] > dims[0] // 4:
                continue
            for j in range(2):
                if not crop[1][j] > dims[1] // 4:
                    continue
                for k in range(2 if sz > 1 else 1):
                    if sz > 1 and not crop[2][k] > dims[2] // 4:
                        continue
                    zs = slice(org[2][k], org[2][k] + crop[2][k]) if sz > 1 else slice(0, 1)
                    sl = (zs, slice(org[1][j], org[1][j] + crop[1][j]), slice(org[0][i], org[0][i] + crop[0][i]))
                    cc = _zncc1(img1[sl], img_t[sl])
                    if best < cc:
                        best, ind = cc, (i, j, k)
        for d in range(3):
            if ind[d] == 1:
                shift[d] = shift[d] - dims[d] if shift[d] > 0 else shift[d] + dims[d]
    return shift


def imshift(vol, shift):
    """imshiftgpukernel, include/cukernel.cuh:476-489: out[x] = in[x - shift], zero outside; shift = (dx, dy, dz)."""
    vol = _f32(vol)
    out = np.zeros_like(vol)
    sz, sy, sx = vol.shape
    dx, dy, dz = (int(v) for v in shift)

    def rng(n, d):
        lo, hi = max(0, d), min(n, n + d)
        return slice(lo, hi), slice(lo - d, hi - d)
    (zo, zi), (yo, yi), (xo, xi) = rng(sz, dz), rng(sy, dy), rng(sx, dx)
    if zo.start < zo.stop and yo.start < yo.stop and xo.start < xo.stop:
        out[zo, yo, xo] = vol[zi, yi, xi]
    return out
