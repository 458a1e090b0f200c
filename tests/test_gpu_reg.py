"""GPU parity of the affine-warp + ZNCC cost and of the whole affine registration against the
CPU oracle (oracle/reg_oracle.*), through the C-ABI.  Tolerances from BASELINE.json north_star:
ZNCC within 1e-5, matrices within 1e-3 voxel-equivalent displacement."""
import numpy as np
import pytest

from microimagelib_b200 import synth

pytestmark = pytest.mark.gpu

IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def _pair(shape=(40, 56, 72), seed=3):
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 2.0))
    tgt = synth.bead_image(shape, psf, seed=seed, density=1 / 2048.0)
    sz, sy, sx = shape
    m = synth.affine_matrix(rot_z_deg=2.0, scale=(1.02, 0.99, 1.0), shift=(1.5, -1.25, 0.75), center=(sx / 2, sy / 2, sz / 2))
    src = synth.warp_exact(tgt, m)
    return tgt, src, m


def _matrices(rng, k, mag=1.0):
    out = []
    for _ in range(k):
        m = IDENT.copy()
        m[[0, 5, 10]] += rng.uniform(-0.03, 0.03, 3).astype(np.float32) * mag
        m[[1, 2, 4, 6, 8, 9]] += rng.uniform(-0.03, 0.03, 6).astype(np.float32) * mag
        m[[3, 7, 11]] += rng.uniform(-3, 3, 3).astype(np.float32) * mag
        out.append(m)
    return np.stack(out)


def corner_disp(m1, m2, shape):
    sz, sy, sx = shape
    d = (np.asarray(m1, np.float64) - np.asarray(m2, np.float64)).reshape(3, 4)
    worst = 0.0
    for x in (0, sx - 1):
        for y in (0, sy - 1):
            for z in (0, sz - 1):
                worst = max(worst, float(np.linalg.norm(d @ np.array([x, y, z, 1.0]))))
    return worst


def test_warp_bit_exact_vs_oracle():
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    tgt, src, m = _pair()
    rng = np.random.default_rng(0)
    for mat in list(_matrices(rng, 3)) + [IDENT, m]:
        got = device.affine_warp(src, mat)
        ref = ro.affine_warp(src, mat)
        assert np.array_equal(got, ref)
    # different output size (resize path) and a big rotation
    big = synth.affine_matrix(rot_z_deg=35.0, scale=(0.7, 1.3, 1.0), shift=(4, -3, 2))
    got = device.affine_warp(src, big, out_shape=(30, 70, 50))
    ref = ro.affine_warp(src, big, out_shape=(30, 70, 50))
    assert np.array_equal(got, ref)


def test_identity_and_integer_shift_kats():
    from microimagelib_b200 import device
    tgt, src, m = _pair()
    got = device.affine_warp(tgt, IDENT)
    assert np.array_equal(got, tgt)             # identity warp is the identity (KAT 3)
    sh = IDENT.copy()
    sh[3], sh[7], sh[11] = 3, -2, 1               # out(x,y,z) = in(x+3, y-2, z+1)
    got = device.affine_warp(tgt, sh)
    want = np.zeros_like(tgt)
    want[:-1, 2:, :-3] = tgt[1:, :-2, 3:]
    assert np.array_equal(got, want)            # integer shift == imshift (KAT 4)


def test_zncc_sums_and_cost_vs_oracle():
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    tgt, src, m = _pair()
    r = device.Reg(tgt.shape)
    r.set_images(tgt, src)
    sd = r.prepare()
    src_dm, _ = ro.demean(src)
    tgt_dm, sd_ref = ro.demean(tgt)
    assert abs(float(sd) - float(sd_ref)) <= 1e-6 * float(sd_ref)
    rng = np.random.default_rng(1)
    mats = np.concatenate([_matrices(rng, 6), IDENT[None], m[None]])
    ss, st = r.cost_sums(mats)
    costs = r.cost(mats)
    for k in range(len(mats)):
        ss_ref, st_ref = ro.zncc_sums(tgt_dm, src_dm, mats[k])
        assert abs(ss[k] - ss_ref) <= 1e-11 * abs(ss_ref)
        assert abs(st[k] - st_ref) <= 1e-11 * max(abs(st_ref), abs(ss_ref) ** 0.5 * float(sd_ref))
        c_ref = ro.cost_from_sums(ss_ref, st_ref, sd_ref)
        assert abs(float(costs[k]) - c_ref) <= 1e-5   # north_star: ZNCC within 1e-5
    # K-batched launches give exactly the values of single launches
    single = np.array([r.cost(mats[k:k + 1])[0] for k in range(len(mats))], np.float32)
    assert np.array_equal(single, costs)
    # identity on identical volumes: ZNCC == 1 (KAT 3)
    r.set_images(tgt, tgt)
    r.prepare()
    assert abs(float(r.cost(IDENT)[0]) + 1.0) < 1e-6
    r.close()


def test_zncc_matches_hardware_texture_filtering():
    """Pins the texture-filter restatement on real hardware: the software fetch used by the product
    and the oracle must agree with a genuine tex3D (linear filter, clamp) to float rounding --
    including the unit's 8-bit rounding of the corner-weight products (oracle/reg_oracle.c)."""
    import ctypes as C
    from microimagelib_b200 import _lib
    lib = _lib.load()
    if not hasattr(lib, "milb_debug_tex3d_warp"):
        pytest.skip("debug texture kernel not built")
    from microimagelib_b200 import device
    tgt, src, m = _pair()
    F = C.POINTER(C.c_float)
    lib.milb_debug_tex3d_warp.argtypes = [F, F, C.POINTER(C.c_uint), F]
    lib.milb_debug_tex3d_warp.restype = C.c_int
    rng = np.random.default_rng(5)
    for mat in list(_matrices(rng, 2)) + [m]:
        hw = np.zeros_like(src)
        size = (C.c_uint * 3)(src.shape[2], src.shape[1], src.shape[0])
        mm = np.ascontiguousarray(mat, np.float32)
        assert lib.milb_debug_tex3d_warp(hw.ctypes.data_as(F), src.ctypes.data_as(F), size, mm.ctypes.data_as(F)) == 0
        sw = device.affine_warp(src, mat)
        scale = float(np.abs(src).max())
        assert float(np.abs(hw - sw).max()) <= 2e-3 * scale
        assert float(np.abs(hw - sw).mean()) <= 1e-4 * scale


@pytest.mark.parametrize("method", [2, 6, 7, 5])
def test_registration_matches_oracle(method):
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    tgt, src, m_true = _pair()
    reg, tmx, st, rec = libapi.reg3d(tgt, src, regChoice=2, regMethod=method, FTOL=1e-4, itLimit=3000)
    assert st == 0
    ref = ro.reg3d_affine(tgt, src, method, ftol=1e-4, it_limit=3000)
    assert abs(float(rec[1]) - float(ref["records"][1])) <= 1e-5
    assert abs(float(rec[3]) - float(ref["records"][3])) <= 1e-5
    assert corner_disp(tmx, ref["tmx"], tgt.shape) <= 1e-3   # north_star tolerance
    assert int(rec[5]) == int(ref["records"][5])
    assert np.array_equal(reg, ref["reg"]) or corner_disp(tmx, ref["tmx"], tgt.shape) > 0
    # and it found the transform we applied (method 2 is rigid only: looser)
    if method != 2:
        assert corner_disp(tmx, synth.invert_affine(m_true), tgt.shape) < 1.5   # src(x) = tgt(M x) => recovers M^-1 (the reference's own run: 1.07 for method 6)
        assert float(rec[3]) > 0.9


@pytest.mark.hw_fetch
def test_hardware_fetch_cost_within_tolerance_of_the_oracle():
    """Default product path: the source is sampled by the texture unit.  Sums within 1e-6 relative of the
    software restatement (<= 2 ulp per sample), costs within the north-star 1e-5."""
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    tgt, src, m = _pair()
    src_dm, _ = ro.demean(src)
    tgt_dm, sd_ref = ro.demean(tgt)
    rng = np.random.default_rng(2)
    mats = np.concatenate([_matrices(rng, 6), IDENT[None], m[None]])
    out = {}
    for mode in ("hw", "sw"):
        r = device.Reg(tgt.shape, fetch=mode)
        r.set_images(tgt, src)
        r.prepare()
        out[mode] = (r.cost_sums(mats), r.cost(mats))
        single = np.array([r.cost(mats[k:k + 1])[0] for k in range(len(mats))], np.float32)
        assert np.array_equal(single, out[mode][1])      # K-batched == single launches in either mode
        r.close()
    (ss_h, st_h), c_h = out["hw"]
    (ss_s, st_s), c_s = out["sw"]
    for k in range(len(mats)):
        ss_ref, st_ref = ro.zncc_sums(tgt_dm, src_dm, mats[k])
        assert abs(ss_s[k] - ss_ref) <= 1e-11 * abs(ss_ref)
        assert abs(ss_h[k] - ss_ref) <= 1e-6 * abs(ss_ref)
        assert abs(st_h[k] - st_ref) <= 1e-6 * max(abs(st_ref), abs(ss_ref) ** 0.5 * float(sd_ref))
        assert abs(float(c_h[k]) - ro.cost_from_sums(ss_ref, st_ref, sd_ref)) <= 1e-5


@pytest.mark.hw_fetch
@pytest.mark.parametrize("method", [6, 7])
def test_registration_hardware_fetch_vs_oracle(method):
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    tgt, src, m_true = _pair()
    reg, tmx, st, rec = libapi.reg3d(tgt, src, regChoice=2, regMethod=method, FTOL=1e-4, itLimit=3000)
    ref = ro.reg3d_affine(tgt, src, method, ftol=1e-4, it_limit=3000)
    assert st == 0
    assert abs(float(rec[1]) - float(ref["records"][1])) <= 1e-5
    assert float(rec[3]) > 0.9 and abs(float(rec[3]) - float(ref["records"][3])) <= 2e-3
    assert corner_disp(tmx, synth.invert_affine(m_true), tgt.shape) < 1.5     # (this trajectory is the reference's own, bit for bit)


def test_registration_with_input_matrix_and_choice0():
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    tgt, src, m_true = _pair()
    guess = synth.affine_matrix(rot_z_deg=1.0, scale=(1, 1, 1), shift=(1, -1, 0.5), center=(36, 28, 20))
    reg, tmx, st, rec = libapi.reg3d(tgt, src, regChoice=2, regMethod=6, inputTmx=True, iTmx=guess)
    ref = ro.reg3d_affine(tgt, src, 6, flag_tmx=True, itmx=guess)
    assert st == 0
    assert corner_disp(tmx, ref["tmx"], tgt.shape) <= 1e-3
    reg0, tmx0, st, _ = libapi.reg3d(tgt, src, regChoice=0, regMethod=6, inputTmx=True, iTmx=guess)
    assert st == 0 and np.array_equal(tmx0, guess)
    assert np.array_equal(reg0, ro.affine_warp(src, guess))
    reg0, tmx0, st, _ = libapi.reg3d(tgt, src, regChoice=0, inputTmx=False)
    assert np.array_equal(reg0, src) and np.array_equal(tmx0, IDENT)


def test_reg3d_error_conventions():
    from microimagelib_b200 import libapi
    tgt, src, _ = _pair((16, 16, 16))
    _, _, st, _ = libapi.reg3d(tgt, src, gpuMemMode=0)
    assert st == -1     # src/api_reg.cpp:390-393
    _, _, st, _ = libapi.reg3d(tgt, src, regChoice=9)
    assert st == 1      # src/api_reg.cpp:512-515
    _, _, st, _ = libapi.reg3d(tgt, src, gpuMemMode=5)
    assert st == 1      # src/api_reg.cpp:591-594


def test_source_size_mismatch_is_centre_aligned():
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do, reg_oracle as ro
    tgt, src, _ = _pair((24, 32, 40))
    small = src[2:-3, 1:-2, 4:-1].copy()
    reg, tmx, st, _ = libapi.reg3d(tgt, small, regChoice=0, inputTmx=False)
    assert st == 0
    assert np.array_equal(reg, do.align_size(small, tgt.shape))


def test_software_fetch_equals_hardware_fetch_on_random_samples():
    """20000 random sub-voxel samples of a white-noise volume: software restatement == tex3D."""
    import ctypes as C
    from microimagelib_b200 import _lib
    lib = _lib.load()
    F = C.POINTER(C.c_float)
    lib.milb_debug_tex3d_sample.argtypes = [F, F, C.POINTER(C.c_uint), F, C.c_int, C.c_int]
    rng = np.random.default_rng(11)
    vol = (rng.random((16, 20, 24)) * 1000).astype(np.float32)
    n = 20000
    # the whole range the kernels sample (0 <= t < size), i.e. including the half texel next to every face where clamp
    # addressing applies, plus exact 8-bit / 9-bit grid points (rounding ties of the coordinate and of the weight products)
    c = np.stack([rng.random(n) * 24, rng.random(n) * 20, rng.random(n) * 16], axis=1).astype(np.float32)
    c[:3000] = np.round(c[:3000] * 256) / 256 + 0.5
    c[3000:6000] = np.round(c[3000:6000] * 512) / 512
    c = np.minimum(c, np.array([24, 20, 16], np.float32) - np.float32(1e-3)).astype(np.float32)
    size = (C.c_uint * 3)(24, 20, 16)
    out = {}
    for hw in (1, 0):
        o = np.zeros(n, np.float32)
        assert lib.milb_debug_tex3d_sample(o.ctypes.data_as(F), vol.ctypes.data_as(F), size, c.ctypes.data_as(F), n, hw) == 0
        out[hw] = o
    assert float(np.abs(out[1] - out[0]).max()) <= 1e-3      # values up to 1000: float rounding only
