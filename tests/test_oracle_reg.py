"""Pins the CPU oracle of the registration path (oracle/reg_oracle.*) with known-answer tests
(SURVEY 8(c) KATs 3-5) and against a float64 numpy restatement of the texture formula."""
import numpy as np

from microimagelib_b200 import synth
from oracle import reg_oracle as ro

IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def _vol(shape=(12, 14, 16), seed=0):
    return np.random.default_rng(seed).random(shape).astype(np.float32) * 100


def test_identity_warp_and_integer_shift():
    v = _vol()
    assert np.array_equal(ro.affine_warp(v, IDENT), v)
    sh = IDENT.copy()
    sh[3], sh[7], sh[11] = 2, -1, 3
    got = ro.affine_warp(v, sh)
    want = np.zeros_like(v)
    want[:-3, 1:, :-2] = v[3:, :-1, 2:]
    assert np.array_equal(got, want)


def _tex_ref(v, tx, ty, tz):
    """The B200 texture unit's linear filter as pinned on the reference's own tex3D output (see oracle/reg_oracle.c):
    xB = t - 0.5 clamped to [0, n - 1], rounded to 8 fractional bits; corner weights from two rounded products with ties
    up for the dx = 1 corners and down for the dx = 0 corners -- integer weights, float64 sum."""
    out = []
    for t, n in ((tx, v.shape[2]), (ty, v.shape[1]), (tz, v.shape[0])):
        xb = min(max(float(np.float32(t) - np.float32(0.5)), 0.0), float(n - 1))
        f = int(np.floor(xb * 256 + 0.5))
        i, a = f >> 8, f & 255
        out.append((min(i, n - 1), min(i + 1, n - 1), a))
    (x0, x1, a), (y0, y1, b), (z0, z1, c) = out
    r = 0.0
    for zz, wz in ((z0, 256 - c), (z1, c)):
        for dx, (xx, wx) in enumerate(((x0, 256 - a), (x1, a))):
            tie = 128 if dx else 127
            wxz = (wx * wz + tie) >> 8
            hi = (wxz * b + tie) >> 8
            r += (wxz - hi) * float(v[zz, y0, xx]) + hi * float(v[zz, y1, xx])
    return r / 256.0


def _coord(m, r, x, y, z):
    """float32 coordinate exactly as the reference build forms it (SASS of oracle/_ref): a1*y, fma a0*x, fma a2*z, add a3, add 0.5"""
    f = np.float32
    a = m[4 * r:4 * r + 4].astype(np.float32)
    t = f(a[1] * f(y))
    t = f(np.float64(a[0]) * np.float64(f(x)) + np.float64(t))
    t = f(np.float64(a[2]) * np.float64(f(z)) + np.float64(t))
    return f(f(t + a[3]) + f(0.5))


def test_fractional_warp_matches_texture_formula():
    v = _vol()
    m = synth.affine_matrix(rot_z_deg=7.0, scale=(1.05, 0.93, 1.1), shift=(0.3, -0.7, 0.45))
    got = ro.affine_warp(v, m)
    M = m.reshape(3, 4).astype(np.float64)
    rng = np.random.default_rng(1)
    for _ in range(300):
        z, y, x = (int(rng.integers(0, s)) for s in v.shape)
        t = [float(ro.lib() and _coord(m, r, x, y, z)) for r in range(3)]
        inside = all(0 <= t[i] < n for i, n in zip(range(3), (v.shape[2], v.shape[1], v.shape[0])))
        want = _tex_ref(v, *t) if inside else 0.0
        assert abs(got[z, y, x] - want) <= 2e-4     # only the float32 accumulation differs


def test_cost_mask_differs_from_warp_mask():
    # corrkernel uses 0 < t (cukernel.cuh:543); affinetransformkernel uses 0 <= t (:514)
    v = np.ones((4, 4, 4), np.float32)
    m = IDENT.copy()
    m[3] = -0.5                     # tx = x - 0.5 + 0.5 = x -> exactly 0 at x = 0
    w = ro.affine_warp(v, m)
    assert w[0, 0, 0] == 1.0        # t = 0 is inside for the warp ...
    ss, st = ro.zncc_sums(v, v, m)
    assert ss == 4 * 4 * 3          # ... and outside for the cost


def test_zncc_identity_is_one_and_cost_conventions():
    v = _vol((10, 12, 14), 3)
    d, sd = ro.demean(v)
    assert abs(ro.zncc_cost(d, sd, d, IDENT) + 1.0) < 1e-6
    far = IDENT.copy()
    far[3] = 1000
    assert ro.zncc_cost(d, sd, d, far) == 2.0     # sum s^2 == 0 -> corrfunc -2 -> cost +2 (:986)


def test_parameter_matrix_maps():
    x = np.arange(13, dtype=np.float32)
    m = ro.p2matrix(x)
    assert list(m) == [4, 5, 6, 1, 7, 8, 9, 2, 10, 11, 12, 3]
    assert np.array_equal(ro.matrix2p(m)[1:], x[1:])
    q = np.array([0, 1, 2, 3, 0, 0, 0, 1, 1, 1], np.float32)
    assert list(ro.dof9tomatrix(q, 3)) == [1, 0, 0, 1, 0, 1, 0, 2, 0, 0, 1, 3]
    q[4] = 57.3 * np.pi / 2        # alpha = 90 degrees about z
    m = ro.dof9tomatrix(q, 6).reshape(3, 4)
    assert np.allclose(m[:, :3], [[0, 1, 0], [-1, 0, 0], [0, 0, 1]], atol=1e-6)   # Rz = [[c, s], [-s, c]]
    q[4] = 0
    q[7:] = [2, 3, 4]
    assert np.allclose(ro.dof9tomatrix(q, 9).reshape(3, 4)[:, :3], np.diag([2, 3, 4]))
    assert np.allclose(ro.dof9tomatrix(q, 7).reshape(3, 4)[:, :3], np.diag([2, 2, 2]))
    a = synth.affine_matrix(3, (1.1, .9, 1), (1, 2, 3))
    b = synth.affine_matrix(-5, (1, 1.2, .8), (-1, 0, 2))
    A = np.vstack([a.reshape(3, 4), [0, 0, 0, 1]]).astype(np.float64)
    B = np.vstack([b.reshape(3, 4), [0, 0, 0, 1]]).astype(np.float64)
    assert np.allclose(ro.matrixmultiply(a, b), (A @ B)[:3].reshape(12), atol=1e-5)


def test_checkmatrix_bounds():
    assert ro.checkmatrix(IDENT, 100, 100, 100)
    bad = IDENT.copy(); bad[0] = 0.4
    assert not ro.checkmatrix(bad, 100, 100, 100)
    bad = IDENT.copy(); bad[3] = 81
    assert not ro.checkmatrix(bad, 100, 100, 100)
    bad = IDENT.copy(); bad[[0, 5, 10]] = 1.39
    assert not ro.checkmatrix(bad, 100, 100, 100)      # scale sum > 4


def test_known_affine_is_recovered_small():
    """KAT 5 at a size the oracle finishes in seconds."""
    psf = synth.gaussian_psf((9, 9, 9), (1.5, 1.5, 1.5))
    tgt = synth.bead_image((24, 32, 40), psf, density=1 / 512.0, seed=5)
    m_true = synth.affine_matrix(rot_z_deg=1.5, scale=(1.01, 0.99, 1.0), shift=(0.8, -0.6, 0.4), center=(20, 16, 12))
    src = synth.warp_exact(tgt, m_true)
    r = ro.reg3d_affine(tgt, src, 6)
    assert r["records"][3] > 0.95 and r["records"][3] > r["records"][1]
    # src(x) = tgt(M x) and the result maps target voxel -> source voxel, so it recovers M^-1
    d = (r["tmx"].astype(np.float64) - synth.invert_affine(m_true)).reshape(3, 4)
    worst = max(np.linalg.norm(d @ np.array([x, y, z, 1.0])) for x in (0, 39) for y in (0, 31) for z in (0, 23))
    assert worst < 1.0
