"""Known-answer tests of the pre-alignment oracle (oracle/reg_oracle.py: phasor, reg2d_shiftalign,
reg2d_affine, max_projection, imshift, tex2d) -- the CPU restatement of reg3d's regChoice 1 / 3 / 4
and of reg2d.  CPU only."""
import numpy as np
import pytest

from microimagelib_b200 import synth
from oracle import reg_oracle as ro


def _vol(shape=(24, 40, 48), seed=5):
    psf = synth.gaussian_psf((9, 9, 9), (1.5, 1.5, 1.5))
    return synth.bead_image(shape, psf, seed=seed, density=1 / 512.0)


def test_tex2d_texel_centres_and_midpoints():
    rng = np.random.default_rng(0)
    img = rng.uniform(0, 100, (9, 11)).astype(np.float32)
    ys, xs = np.mgrid[0:9, 0:11]
    c = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], 1).astype(np.float32)
    assert np.array_equal(ro.tex2d_samples(img, c), img.ravel())          # texel centres are exact
    mid = ro.tex2d_samples(img, [[3.0, 4.5]])[0]                           # halfway between x = 2 and x = 3 on row 4
    # integer weights 128 + 128, exact sum, ONE rounding to float with ties away from zero (what the texture unit does), exact 1/256
    s = 128 * np.float64(img[4, 2]) + 128 * np.float64(img[4, 3])              # exact in double
    f = np.float32(s)
    if np.float64(f) != s:
        g = np.nextafter(f, np.float32(np.inf) if s > np.float64(f) else np.float32(-np.inf))
        if abs(np.float64(g) - s) == abs(np.float64(f) - s) and abs(g) > abs(f):
            f = g                                                               # an exact tie: away from zero
    assert mid == np.float32(f / np.float32(256))
    edge = ro.tex2d_samples(img, [[0.1, 0.2], [10.9, 8.9]])               # clamp addressing outside the first / last centre
    assert edge[0] == img[0, 0] and edge[1] == img[8, 10]


def test_imshift_matches_definition():
    v = _vol()
    out = ro.imshift(v, (3, -2, 1))                                        # out[z, y, x] = v[z - 1, y + 2, x - 3]
    want = np.zeros_like(v)
    want[1:, :-2, 3:] = v[:-1, 2:, :-3]
    assert np.array_equal(out, want)
    assert np.array_equal(ro.imshift(v, (0, 0, 0)), v)
    assert not ro.imshift(v, (100, 0, 0)).any()


def test_max_projection_layouts():
    v = _vol((5, 6, 7))
    assert ro.max_projection(v, 1).shape == (6, 7)
    p2 = ro.max_projection(v, 2)
    assert p2.shape == (7, 5) and p2[3, 2] == v[2, :, 3].max()             # rows x, columns z
    p3 = ro.max_projection(v, 3)
    assert p3.shape == (5, 6) and p3[4, 1] == v[4, 1, :].max()
    assert ro.max_projection(-np.abs(v) - 1, 1).max() == 0                 # the running maximum starts at 0


@pytest.mark.parametrize("shift", [(3, -2, 1), (0, 0, 0), (-5, 4, -3)])
def test_phasor_recovers_circular_shift(shift):
    v = _vol()
    dx, dy, dz = shift
    moved = np.roll(v, (dz, dy, dx), axis=(0, 1, 2))                       # moved(x) = v(x - d)
    assert ro.phasor(v, moved) == [dx, dy, dz]


def test_phasor_large_shift_picks_the_overlapping_alias():
    # a true (non-circular) shift of more than a quarter of the extent: the correlation peak is
    # ambiguous modulo the size, and the ZNCC of the overlap decides (src/api_subfunc.cu:2497-2587)
    v = _vol((24, 40, 96), seed=11)
    moved = ro.imshift(v, (30, 0, 0))
    assert ro.phasor(v, moved) == [30, 0, 0]
    moved = ro.imshift(v, (-30, 0, 0))
    assert ro.phasor(v, moved) == [-30, 0, 0]


def test_phasor_2d_and_odd_sizes():
    v = ro.max_projection(_vol((12, 37, 45), seed=2), 1)
    moved = np.roll(v, (4, -6), axis=(0, 1))
    assert ro.phasor(v, moved) == [-6, 4, 0]


def test_shiftalign_recovers_grid_shift():
    img = ro.max_projection(_vol((16, 60, 80), seed=3), 1)
    # source(x) = img(x - (4, -3)): warping the source by +(4, -3) restores the target; the search grid
    # has pitch 80*0.3/30 = 0.8 px in x and 60*0.3/30 = 0.6 px in y, so 4.0 = 5 steps, -3.0 = -5 steps
    src = ro.imshift(img[None], (4, -3, 0))[0]
    r = ro.reg2d_shiftalign(img, src, search_y=True, shift_region=0.3, total_step=30.0)
    assert abs(r["tmx"][2] - 4.0) < 1e-4 and abs(r["tmx"][5] + 3.0) < 1e-4
    assert r["best"] > r["initial"] and r["best"] > 0.95
    # x-only variant keeps the y entry of the input matrix
    r2 = ro.reg2d_shiftalign(img, src, flag_tmx=True, itmx=[1, 0, 0, 0, 1, -3], search_y=False)
    assert abs(r2["tmx"][2] - 4.0) < 1e-4 and r2["tmx"][5] == -3


def test_shiftalign_no_positive_candidate_returns_zero_shift():
    a = ro.max_projection(_vol((12, 40, 40), seed=4), 1)                   # smooth: small shifts stay anticorrelated
    r = ro.reg2d_shiftalign(a, -a, search_y=True, shift_region=0.05, total_step=2.0)   # anticorrelated everywhere nearby
    assert r["tmx"][2] == 0 and r["tmx"][5] == 0


def test_reg2d_affine_improves_and_is_deterministic():
    img = ro.max_projection(_vol((16, 48, 64), seed=8), 1)
    src = ro.affine2d(img, [1.01, 0.01, -1.5, -0.01, 0.99, 1.0], img.shape)
    r1 = ro.reg2d_affine(img, src, ftol=1e-4, it_limit=400)
    r2 = ro.reg2d_affine(img, src, ftol=1e-4, it_limit=400)
    assert r1["best"] > r1["initial"] and r1["best"] > 0.95
    assert np.array_equal(r1["tmx"], r2["tmx"]) and r1["n_eval"] == r2["n_eval"]


def test_prealign_mip_translation():
    v = _vol((32, 48, 64), seed=9)
    src = ro.imshift(v, (3.0, -2.0, 2.0))
    m = ro.prealign_mip(v, src)
    # grid pitches: x 0.64, y 0.48 (XY MIP), z 0.32 (ZX MIP): the nearest grid points to (3, -2, 2)
    assert abs(m[3] - 3.0) <= 0.33 and abs(m[7] + 2.0) <= 0.25 and abs(m[11] - 2.0) <= 0.17
    assert m[0] == 1 and m[5] == 1 and m[10] == 1
