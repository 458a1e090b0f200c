"""GPU parity of the pre-alignment rows (reg3d regChoice 1 / 3 / 4, reg2d) against the CPU oracle,
through the C-ABI: hardware tex2D == software tex2D == oracle, 2-D cost within 1e-6, shift-search
matrices and phasor shifts identical, libapi reg3d / reg2d outputs equal to the oracle chain."""
import numpy as np
import pytest

from microimagelib_b200 import synth

pytestmark = pytest.mark.gpu


def _vol(shape=(24, 40, 48), seed=5):
    psf = synth.gaussian_psf((9, 9, 9), (1.5, 1.5, 1.5))
    return synth.bead_image(shape, psf, seed=seed, density=1 / 512.0)


def test_tex2d_hardware_equals_software_equals_oracle():
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    rng = np.random.default_rng(1)
    img = rng.uniform(-50, 200, (37, 53)).astype(np.float32)
    n = 40000
    c = np.stack([rng.uniform(-1.0, 54.0, n), rng.uniform(-1.0, 38.0, n)], 1).astype(np.float32)
    c[:2000] = np.round(c[:2000] * 256) / 256 + 0.5          # exact 8-bit fractions, incl. rounding ties of the products
    c[2000:4000] = np.round(c[2000:4000] * 512) / 512          # half-way points of the 8-bit coordinate grid
    hw = device.tex2d_samples(img, c, hardware=True)
    sw = device.tex2d_samples(img, c, hardware=False)
    orc = ro.tex2d_samples(img, c)
    assert np.array_equal(sw, orc)
    # the unit's integer corner weights are reproduced exactly (scripts/tex_probe2d.py: all 65536 fraction
    # pairs); its final float accumulation differs from the canonical fma chain by at most an ulp or two
    assert float(np.abs(hw - sw).max()) <= 4e-5, f"{np.count_nonzero(hw != sw)} of {n} samples differ, max {np.abs(hw - sw).max()}"


def test_reg2d_cost_matches_oracle():
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    img = ro.max_projection(_vol((16, 60, 80), seed=3), 1)
    src = ro.imshift(img[None], (4, -3, 0))[0] + np.float32(5)
    rng = np.random.default_rng(2)
    mats = np.tile(np.array([1, 0, 0, 0, 1, 0], np.float32), (40, 1))
    mats[:, [0, 4]] += rng.uniform(-0.05, 0.05, (40, 2)).astype(np.float32)
    mats[:, [1, 3]] += rng.uniform(-0.05, 0.05, (40, 2)).astype(np.float32)
    mats[:, [2, 5]] += rng.uniform(-8, 8, (40, 2)).astype(np.float32)
    mats[-1] = [1, 0, 500, 0, 1, 0]                                         # everything outside: cost +2
    r = device.Reg2D(img, src)
    tgt_dm, sd_t = ro.demean(img)
    src_dm, _ = ro.demean(src)
    assert r.sd_t == sd_t
    got = r.cost(mats)
    want = ro.corr2d_costs(tgt_dm, sd_t, src_dm, mats)
    assert got[-1] == 2.0 and want[-1] == 2.0
    assert np.abs(got - want).max() <= 1e-6
    for m in mats[:3]:                                                       # warps are bit-exact
        assert np.array_equal(r.warp(m, raw_source=True), ro.affine2d(src, m, img.shape))
    # one candidate at a time (many blocks per candidate) gives the same numbers as the batched launch
    one = np.array([r.cost(m[None])[0] for m in mats[:5]])
    assert np.abs(one - got[:5]).max() <= 1e-7
    r.close()


@pytest.mark.parametrize("search_y", [True, False])
def test_shiftalign_matches_oracle(search_y):
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    img = ro.max_projection(_vol((16, 60, 80), seed=3), 1)
    src = ro.imshift(img[None], (4, -3, 0))[0]
    itmx = None if search_y else [1, 0, 0, 0, 1, -3]
    reg, tmx, rec = device.reg2d_shiftalign(img, src, flag_tmx=not search_y, itmx=itmx, search_y=search_y)
    o = ro.reg2d_shiftalign(img, src, flag_tmx=not search_y, itmx=itmx, search_y=search_y)
    assert np.array_equal(tmx, o["tmx"])
    assert abs(rec[4] - o["initial"]) <= 1e-6 and abs(rec[5] - o["best"]) <= 1e-6
    assert np.array_equal(reg, o["reg"])


def test_reg2d_affine_matches_oracle():
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    img = ro.max_projection(_vol((16, 48, 64), seed=8), 1)
    src = ro.affine2d(img, [1.01, 0.01, -1.5, -0.01, 0.99, 1.0], img.shape)
    reg, tmx, rec = device.reg2d_affine(img, src, ftol=1e-4, it_limit=400)
    o = ro.reg2d_affine(img, src, ftol=1e-4, it_limit=400)
    assert int(rec[5]) == o["n_eval"]
    assert np.abs(tmx - o["tmx"]).max() <= 1e-3
    assert abs(rec[3] - o["best"]) <= 1e-5 and abs(rec[1] - o["initial"]) <= 1e-5


@pytest.mark.parametrize("shape,shift", [((24, 40, 48), (3, -2, 1)), ((24, 40, 48), (0, 0, 0)), ((21, 37, 45), (-5, 4, -3)),
                                          ((50, 36, 100), (7, 0, -9))])
def test_phasor_circular_shift(shape, shift):
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    v = _vol(shape)
    dx, dy, dz = shift
    moved = np.roll(v, (dz, dy, dx), axis=(0, 1, 2))
    assert device.phasor(v, moved) == ro.phasor(v, moved) == [dx, dy, dz]


def test_phasor_large_shift_alias_and_2d():
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    v = _vol((24, 40, 96), seed=11)
    for d in ((30, 0, 0), (-30, 0, 0), (28, -12, 7)):
        moved = ro.imshift(v, d)
        assert device.phasor(v, moved) == ro.phasor(v, moved)
    assert device.phasor(v, ro.imshift(v, (30, 0, 0))) == [30, 0, 0]
    img = ro.max_projection(_vol((12, 37, 45), seed=2), 1)
    moved = np.roll(img, (4, -6), axis=(0, 1))
    assert device.phasor(img, moved) == ro.phasor(img, moved) == [-6, 4, 0]
    big = ro.imshift(img[None], (20, -15, 0))[0]
    assert device.phasor(img, big) == ro.phasor(img, big)


@pytest.mark.parametrize("shape", [(64, 128, 192), (128, 64, 320), (64, 64, 64)])
def test_phasor_on_the_projects_own_fft(shape, monkeypatch):
    """Image sizes the project's 3-D transform handles without padding (multiples of 64 that snapTransformSize keeps) run the
    phase correlation on it (power-of-two and 64*k fast plans); the cuFFT path on the exact size and the CPU oracle give the
    same shifts, including large shifts that need the alias disambiguation."""
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    v = _vol(shape, seed=5)
    for d in ((5, -3, 2), (0, 0, 0), (shape[2] // 2 - 6, -7, 9), (-11, shape[1] // 3, -shape[0] // 3)):
        moved = ro.imshift(v, d)
        want = ro.phasor(v, moved)
        monkeypatch.setenv("MILB_PHASOR_CUFFT", "0")
        l0 = device.launch_count()
        own = device.phasor(v, moved)
        own_launches = device.launch_count() - l0
        monkeypatch.setenv("MILB_PHASOR_CUFFT", "1")
        l0 = device.launch_count()
        lib = device.phasor(v, moved)
        lib_launches = device.launch_count() - l0
        assert own == lib == want
        assert own_launches > lib_launches + 8          # the transform's own kernels were launched (cuFFT's are not counted)
    moved = np.roll(v, (4, -9, 17), axis=(0, 1, 2))
    monkeypatch.setenv("MILB_PHASOR_CUFFT", "0")
    assert device.phasor(v, moved) == [17, -9, 4]


def test_imshift_exact():
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    v = _vol()
    for d in ((3, -2, 1), (0, 0, 0), (-7, 11, -4), (100, 0, 0)):
        assert np.array_equal(device.imshift(v, d), ro.imshift(v, d))


def _shifted_pair(shape=(32, 48, 64), shift=(3.0, -2.0, 2.0), seed=9):
    v = _vol(shape, seed=seed)
    from oracle import reg_oracle as ro
    return v, ro.imshift(v, shift)


def test_libapi_reg3d_phasor_choices():
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    tgt, src = _shifted_pair()
    # regChoice 1: integer shift only
    reg, tmx, st, rec = libapi.reg3d(tgt, src, regChoice=1, regMethod=2)
    sh = ro.phasor(tgt, src)
    assert st == 0 and sh == [3, -2, 2]
    want = np.array([1, 0, 0, sh[0], 0, 1, 0, sh[1], 0, 0, 1, sh[2]], np.float32)
    assert np.array_equal(tmx, want)
    assert np.array_equal(reg, ro.imshift(src, [-s for s in sh]))
    # regChoice 3: phasor shift as the input matrix of the affine registration
    reg3, tmx3, st3, rec3 = libapi.reg3d(tgt, src, regChoice=3, regMethod=2, FTOL=1e-3, itLimit=200)
    o = ro.reg3d_affine(tgt, src, 2, flag_tmx=True, itmx=want, ftol=1e-3, it_limit=200)
    assert st3 == 0 and int(rec3[5]) == int(o["records"][5])
    assert np.abs(tmx3 - o["tmx"]).max() <= 1e-3
    assert abs(rec3[3] - o["records"][3]) <= 1e-5


def test_libapi_reg3d_mip_prealignment():
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    tgt, src = _shifted_pair()
    reg, tmx, st, rec = libapi.reg3d(tgt, src, regChoice=4, regMethod=1, FTOL=1e-3, itLimit=200)
    pre = ro.prealign_mip(tgt, src)
    o = ro.reg3d_affine(tgt, src, 1, flag_tmx=True, itmx=pre, ftol=1e-3, it_limit=200)
    assert st == 0 and int(rec[5]) == int(o["records"][5])
    assert np.abs(tmx - o["tmx"]).max() <= 1e-3
    assert abs(rec[3] - o["records"][3]) <= 1e-5
    assert np.abs(reg - o["reg"]).max() <= 1e-3 * max(1.0, float(np.abs(o["reg"]).max()))
    assert rec[3] > 0.98                                                     # and it did align the pair


def test_libapi_reg2d_choices():
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    img = ro.max_projection(_vol((16, 60, 80), seed=3), 1)
    src = ro.imshift(img[None], (4, -3, 0))[0]
    reg, tmx, st, rec = libapi.reg2d(img, src, regChoice=1)                  # shift search, region 0.4 / 40 steps
    o = ro.reg2d_shiftalign(img, src, search_y=True, shift_region=0.4, total_step=40.0)
    assert st == 0 and np.array_equal(tmx, o["tmx"]) and np.array_equal(reg, o["reg"])
    reg, tmx, st, rec = libapi.reg2d(img, src, regChoice=3)                  # phasor
    assert st == 0 and list(tmx) == [1, 0, 4, 0, 1, -3]
    assert np.array_equal(reg, ro.imshift(src[None], (-4, 3, 0))[0])
    reg, tmx, st, rec = libapi.reg2d(img, src, regChoice=2, FTOL=1e-3, itLimit=300)   # 6-DOF affine
    o = ro.reg2d_affine(img, src, ftol=1e-3, it_limit=300)
    assert st == 0 and int(rec[5]) == o["n_eval"] and np.abs(tmx - o["tmx"]).max() <= 1e-3
    reg, tmx, st, rec = libapi.reg2d(img, src, regChoice=0, flagTmx=True, iTmx=[1, 0, 4, 0, 1, -3])
    assert st == 0 and np.array_equal(reg, ro.affine2d(src, [1, 0, 4, 0, 1, -3], img.shape))


def test_phasor_peak_in_the_first_column_reports_z_zero_like_max3dgpu():
    """max3Dgpu never reads the z index of column (0, 0) of the shifted correlation volume (src/api_subfunc.cu:453-466):
    a shift of exactly (-sx/2, -sy/2, dz) therefore comes back with z = -sz/2.  Oracle and CUDA restate the quirk."""
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    v = _vol((16, 24, 32), seed=21)
    moved = np.roll(v, (3, -12, -16), axis=(0, 1, 2))           # (dz, dy, dx) = (3, -sy/2, -sx/2)
    got, want = device.phasor(v, moved), ro.phasor(v, moved)
    assert got == want
    assert got[0] in (-16, 16) and got[1] in (-12, 12)            # the alias comparison may flip the half-extent shifts


def test_libapi_registration_mode_and_choice_errors():
    """return codes of the reference for unsupported combinations (src/api_reg.cpp:390-393, 514, 583-586, 596-599; :220-236)"""
    from microimagelib_b200 import libapi
    v = _vol((16, 24, 32), seed=22)
    assert libapi.reg3d(v, v, regChoice=2, regMethod=1, gpuMemMode=0)[2] == -1      # CPU registration "under developing"
    assert libapi.reg3d(v, v, regChoice=2, regMethod=1, gpuMemMode=3)[2] == 1       # wrong gpuMemMode
    assert libapi.reg3d(v, v, regChoice=7, regMethod=1)[2] == 1                     # wrong registration choice
    assert libapi.reg3d(v, v, regChoice=4, regMethod=1, gpuMemMode=2)[2] == -1      # 2-D MIP pre-alignment not in mode 2
    img = v[0]
    assert libapi.reg2d(img, img, regChoice=9)[2] == 1
    assert libapi.reg2d(img, img[:, :-2].copy(), regChoice=3)[2] == 1               # phasor needs equal sizes
    reg, tmx, st, rec = libapi.reg3d(v, v, regChoice=0, regMethod=3, inputTmx=False) # choice 0 without a matrix: copy
    assert st == 0 and np.array_equal(reg, v)
