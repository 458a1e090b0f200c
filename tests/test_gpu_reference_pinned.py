"""Pins the oracle AND the product to the reference itself, run on the same B200.

oracle/_ref/libapi_ref.so is the reference's own libapi (src/api_decon.cpp, api_reg.cpp, api_subfunc.cu,
apifunc.cpp, api_powell.c + include/cukernel.cuh) built by oracle/build_ref_gpu.py with a mechanical
texture-object patch (CUDA 12 removed texture references) against cuFFT.  Every test runs the same
inputs through   reference (ref)  /  CPU oracle (orc)  /  product (got)   and checks both
    orc <-> ref   (the oracle restates the reference)   and   got <-> ref   (the product matches it)
at the north-star tolerances: deconvolution rel-L2 <= 1e-4, ZNCC <= 1e-5, matrices <= 1e-3
voxel-equivalent displacement, integer / index work bit-exact.
"""
import numpy as np
import pytest

from microimagelib_b200 import synth
from oracle import ref_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref/libapi_ref.so not built")]

IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def corner_disp(m1, m2, shape):
    sz, sy, sx = shape
    d = (np.asarray(m1, np.float64) - np.asarray(m2, np.float64)).reshape(3, 4)
    return max(float(np.linalg.norm(d @ np.array([x, y, z, 1.0]))) for x in (0, sx - 1) for y in (0, sy - 1) for z in (0, sz - 1))


def _views(shape, psf_shape=(17, 17, 17), seed=20260):
    psf_a = synth.gaussian_psf(psf_shape, (2.0, 2.0, 3.0))
    psf_b = synth.gaussian_psf(psf_shape, (3.0, 2.0, 2.0))
    a = synth.bead_image(shape, psf_a, density=1 / 2048.0, seed=seed)
    b = synth.bead_image(shape, psf_b, density=1 / 2048.0, seed=seed, noise_seed=seed + 2)
    return a, b, psf_a, psf_b


# ---------------------------------------------------------------------------------------- deconvolution
DECON_CASES = [
    # shape (S, H, W), psf shape, iterations, const initial, unmatched back projector
    ((128, 128, 128), (17, 17, 17), 10, False, False),    # BASELINE config 1 (power-of-two box: fused fast kernels)
    ((40, 100, 72), (17, 17, 17), 6, False, False),        # padded box 48 x 128 x 80 (generic path)
    ((64, 64, 64), (16, 16, 16), 5, False, False),         # even PSF: the flip's one-voxel shift (cukernel.cuh:675)
    ((32, 64, 64), (9, 9, 9), 4, True, False),             # constant initial estimate = sum, not mean
    ((64, 64, 64), (17, 17, 17), 5, False, True),          # unmatched back projector
    ((24, 40, 56), (33, 49, 65), 3, False, False),         # PSF larger than the box (alignsize3D branch)
]


@pytest.mark.parametrize("shape,pshape,iters,const_init,unmatch", DECON_CASES)
def test_decon_singleview_three_way(shape, pshape, iters, const_init, unmatch):
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    a, _, psf, psf_b = _views(shape, pshape)
    bp = psf_b if unmatch else None
    ref, st_r, _ = ref_gpu.api().decon_singleview(a, psf, iters, initialFlag=const_init, flagUnmatch=unmatch, psf_bp=bp)
    got, st_g, _ = libapi.decon_singleview(a, psf, iters, initialFlag=const_init, flagUnmatch=unmatch, psf_bp=bp)
    orc = do.decon_singleview(a, psf, iters, const_init=const_init, unmatch=unmatch, psf_bp=bp)
    assert st_r == 0 and st_g == 0
    assert rel_l2(orc, ref) <= 1e-4, "oracle does not restate the reference"
    assert rel_l2(got, ref) <= 1e-4, "product does not match the reference"


@pytest.mark.parametrize("shape,pshape,iters,const_init,unmatch", DECON_CASES[:5])
def test_decon_dualview_three_way(shape, pshape, iters, const_init, unmatch):
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    a, b, psf_a, psf_b = _views(shape, pshape)
    kw = dict(flagUnmatch=unmatch, psf_bp1=psf_b if unmatch else None, psf_bp2=psf_a if unmatch else None)
    ref, st_r, _ = ref_gpu.api().decon_dualview(a, b, psf_a, psf_b, iters, initialFlag=const_init, **kw)
    got, st_g, _ = libapi.decon_dualview(a, b, psf_a, psf_b, iters, initialFlag=const_init, **kw)
    orc = do.decon_dualview(a, b, psf_a, psf_b, iters, const_init=const_init, unmatch=unmatch, psf_bp1=kw["psf_bp1"], psf_bp2=kw["psf_bp2"])
    assert st_r == 0 and st_g == 0
    assert rel_l2(orc, ref) <= 1e-4
    assert rel_l2(got, ref) <= 1e-4


def test_decon_at_the_benchmarked_size_against_the_reference():
    """BASELINE config 2 box (512 x 512 x 256), 10 iterations: product vs the reference's cuFFT loop."""
    from microimagelib_b200 import libapi
    psf = synth.gaussian_psf((33, 33, 33), (4.0, 2.0, 2.0))
    a = synth.bead_image((256, 512, 512), psf, density=1 / 16384.0)
    ref, st_r, _ = ref_gpu.api().decon_singleview(a, psf, 10)
    got, st_g, _ = libapi.decon_singleview(a, psf, 10)
    assert st_r == 0 and st_g == 0
    assert rel_l2(got, ref) <= 1e-4


# ---------------------------------------------------------------------------------------- warp and ZNCC
def _pair(shape=(40, 56, 72), seed=3):
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 2.0))
    tgt = synth.bead_image(shape, psf, seed=seed, density=1 / 2048.0)
    sz, sy, sx = shape
    m = synth.affine_matrix(rot_z_deg=2.0, scale=(1.02, 0.99, 1.0), shift=(1.5, -1.25, 0.75), center=(sx / 2, sy / 2, sz / 2))
    src = synth.warp_exact(tgt, m)
    return tgt, src, m


def test_affine_warp_three_way():
    """atrans3dgpu: the reference samples with the hardware texture unit, and so does the product by default.  (1) The
    hardware fetch at the PRODUCT's coordinates reproduces the reference's output bit for bit (the coordinate expression is
    pinned to the reference build's SASS).  (2) The product's software twin (MILB_TEX_FETCH=sw, tests/test_gpu_reg.py) is
    bit-identical to the oracle.
    (3) Software restatement vs hardware: the integer weights agree on every sample (also in the clamp region next to
    the faces); only the float accumulation order differs, by a few ulp."""
    import ctypes as C
    from microimagelib_b200 import _lib, libapi
    from oracle import reg_oracle as ro
    lib = _lib.load()
    F = C.POINTER(C.c_float)
    lib.milb_debug_tex3d_warp.argtypes = [F, F, C.POINTER(C.c_uint), F]
    tgt, src, m = _pair()
    big = synth.affine_matrix(rot_z_deg=35.0, scale=(0.7, 1.3, 1.0), shift=(4, -3, 2))
    for mat, oshape in ((m, None), (IDENT, None), (big, None), (big, (30, 70, 50))):
        ref, st = ref_gpu.api().atrans3dgpu(src, mat, out_shape=oshape)
        got, st2 = libapi.atrans3dgpu(src, mat, out_shape=oshape)
        orc = ro.affine_warp(src, mat, out_shape=oshape)
        assert st == 0 and st2 == 0
        assert np.array_equal(got, ref)                               # product (hardware fetch, default) == reference, bit for bit
        if oshape is None:
            hw = np.zeros_like(src)
            size = (C.c_uint * 3)(src.shape[2], src.shape[1], src.shape[0])
            mm = np.ascontiguousarray(mat, np.float32)
            assert lib.milb_debug_tex3d_warp(hw.ctypes.data_as(F), src.ctypes.data_as(F), size, mm.ctypes.data_as(F)) == 0
            assert np.array_equal(hw, ref)
        scale = float(np.abs(src).max())
        d = np.abs(ref - orc)
        assert np.array_equal(ref == 0, orc == 0)                      # same validity mask
        assert float(d.max()) <= 4e-6 * scale                          # a few ulp of the largest filtered value


def test_affine_warp_16bit_nearest_three_way():
    """atrans3dgpu_16bit samples tex16, which the reference leaves at point filtering (api_subfunc.cu:909-919)."""
    from microimagelib_b200 import libapi
    tgt, src, m = _pair()
    s16 = np.clip(src, 0, 65535).astype(np.uint16)
    ref, st = ref_gpu.api().atrans3dgpu_16bit(s16, m)
    got, st2 = libapi.atrans3dgpu_16bit(s16, m)
    assert st == 0 and st2 == 0
    assert float(np.mean(ref != got)) <= 1e-4      # nearest-texel ties at exact half-voxel positions only


@pytest.mark.parametrize("method", [1, 2, 3, 4, 5, 6, 7])
def test_reg3d_three_way(method):
    """Whole reg3d (ZNCC cost + Powell) for every affMethod.
    product (default: hardware fetch) vs reference: the SAME trajectory -- identical evaluation count, initial / final
    ZNCC within 1e-5, matrix within 1e-3 voxel-equivalent displacement (measured: bit-identical).
    oracle (software restatement of the fetch) vs reference: every single cost evaluation agrees to ~1e-7, so the
    initial ZNCC is within 1e-5; Powell then amplifies one-ulp cost differences into different line-search brackets, so
    the end points are only equally good optima: final ZNCC within 2e-3, matrices within one voxel."""
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    tgt, src, m_true = _pair()
    _, tmx_r, st_r, rec_r = ref_gpu.api().reg3d(tgt, src, regChoice=2, regMethod=method, FTOL=1e-4, itLimit=3000)
    _, tmx_g, st_g, rec_g = libapi.reg3d(tgt, src, regChoice=2, regMethod=method, FTOL=1e-4, itLimit=3000)
    orc = ro.reg3d_affine(tgt, src, method, ftol=1e-4, it_limit=3000)
    assert st_r == 0 and st_g == 0
    assert abs(float(rec_g[1]) - float(rec_r[1])) <= 1e-5
    assert abs(float(rec_g[3]) - float(rec_r[3])) <= 1e-5
    assert int(rec_g[5]) == int(rec_r[5])
    assert corner_disp(tmx_g, tmx_r, tgt.shape) <= 1e-3
    assert abs(float(orc["records"][1]) - float(rec_r[1])) <= 1e-5
    assert abs(float(orc["records"][3]) - float(rec_r[3])) <= 2e-3
    assert corner_disp(orc["tmx"], tmx_r, tgt.shape) <= 1.0


def test_zncc_at_given_matrices_three_way():
    """records[1] of an affMethod-5 call with an input matrix is corrfunc at exactly that matrix
    (src/api_subfunc.cu:2817-2821, 2881): a direct probe of the reference's ZNCC kernel."""
    from microimagelib_b200 import device, libapi
    from oracle import reg_oracle as ro
    tgt, src, m = _pair()
    rng = np.random.default_rng(5)
    t_dm, sd = ro.demean(tgt)
    s_dm, _ = ro.demean(src)
    for k in range(4):
        mat = IDENT.copy()
        mat[[0, 5, 10]] += rng.uniform(-0.03, 0.03, 3).astype(np.float32)
        mat[[1, 2, 4, 6, 8, 9]] += rng.uniform(-0.03, 0.03, 6).astype(np.float32)
        mat[[3, 7, 11]] += rng.uniform(-3, 3, 3).astype(np.float32)
        _, _, st, rec_r = ref_gpu.api().reg3d(tgt, src, regChoice=2, regMethod=5, inputTmx=True, iTmx=mat, itLimit=1)
        assert st == 0
        orc = -float(ro.zncc_cost(t_dm, sd, s_dm, mat))
        r = device.Reg(tgt.shape)
        r.set_images(tgt, src)
        r.prepare()
        got = -float(r.cost(mat)[0])
        r.close()
        assert abs(orc - float(rec_r[1])) <= 1e-5
        assert abs(got - float(rec_r[1])) <= 1e-5


def test_zncc_config4_size_three_way():
    """One ZNCC evaluation at BASELINE config 4's size (512 x 512 x 256) against the reference's corrkernel."""
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    shape = (256, 512, 512)
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 2.0))
    tgt = synth.bead_image(shape, psf, density=1 / 16384.0)
    m = synth.affine_matrix(rot_z_deg=2.0, scale=(1.02, 0.99, 1.0), shift=(3.5, -2.25, 1.75), center=(256, 256, 128))
    src = ro.affine_warp(tgt, m)
    mat = IDENT.copy()
    mat[3], mat[7], mat[11] = 0.5, -0.25, 0.125
    _, _, st, rec_r = ref_gpu.api().reg3d(tgt, src, regChoice=2, regMethod=5, inputTmx=True, iTmx=mat, itLimit=1)
    assert st == 0
    t_dm, sd = ro.demean(tgt)
    s_dm, _ = ro.demean(src)
    orc = -float(ro.zncc_cost(t_dm, sd, s_dm, mat))
    r = device.Reg(shape)
    r.set_images(tgt, src)
    r.prepare()
    got = -float(r.cost(mat)[0])
    r.close()
    assert abs(orc - float(rec_r[1])) <= 1e-5
    assert abs(got - float(rec_r[1])) <= 1e-5


@pytest.mark.parametrize("choice", [1, 3, 4])
def test_reg3d_prealignment_choices_three_way(choice):
    """regChoice 1 (phasor only), 3 (phasor + affine), 4 (2-D MIP shift search + affine)."""
    from microimagelib_b200 import libapi
    tgt, _, _ = _pair((32, 48, 64))
    from oracle import reg_oracle as ro
    src = ro.imshift(tgt, (3, -2, 1))
    _, tmx_r, st_r, rec_r = ref_gpu.api().reg3d(tgt, src, regChoice=choice, regMethod=6, FTOL=1e-4, itLimit=3000)
    _, tmx_g, st_g, rec_g = libapi.reg3d(tgt, src, regChoice=choice, regMethod=6, FTOL=1e-4, itLimit=3000)
    assert st_r == st_g
    assert corner_disp(tmx_g, tmx_r, tgt.shape) <= 1e-3 or abs(float(rec_g[3]) - float(rec_r[3])) <= 1e-5


# ---------------------------------------------------------------------------------------- geometry, MIPs, files
def test_geometry_and_mips_bit_exact_against_the_reference():
    from microimagelib_b200 import libapi
    rng = np.random.default_rng(9)
    vol = (rng.random((20, 28, 36)) * 4000).astype(np.float32)
    R = ref_gpu.api()
    for op in (1, 2):
        a, _ = R.imoperation3D(vol, op)
        b, _ = libapi.imoperation3D(vol, op)
        assert a.shape == b.shape and np.array_equal(a, b)
    a, _ = R.alignsize3d(vol, (24, 20, 40))
    b, _ = libapi.alignsize3d(vol, (24, 20, 40))
    assert np.array_equal(a, b)
    za, xa, ya, _ = R.mp2dgpu(vol)
    zb, xb, yb, _ = libapi.mp2dgpu(vol)
    assert np.array_equal(za, zb) and np.array_equal(xa, xb) and np.array_equal(ya, yb)
    za, xa, ya, _ = R.mp2dgpu(vol, flagZProj=False)        # the flagZProj gate also switches the Y projection off
    zb, xb, yb, _ = libapi.mp2dgpu(vol, flagZProj=False)
    assert np.array_equal(za, zb) and np.array_equal(xa, xb) and np.array_equal(ya, yb)
    # resampling and rotating projections go through the texture unit in the reference and in the product: identical
    a, _ = R.imresize3d(vol, (30, 28, 36))
    b, _ = libapi.imresize3d(vol, (30, 28, 36))
    assert np.array_equal(a, b)
    for axis in (1, 2):
        a, _ = R.mip3dgpu(vol, axis, 6)
        b, _ = libapi.mip3dgpu(vol, axis, 6)
        assert a.shape == b.shape and np.array_equal(a, b)


def test_tiff_files_interchange_with_the_reference(tmp_path):
    """16-bit and float TIFF stacks: each library reads what the other wrote, identical voxels; the
    16-bit conversion is the C truncation (uint16) in both."""
    from microimagelib_b200 import libapi
    rng = np.random.default_rng(4)
    vol = (rng.random((5, 12, 20)) * 60000).astype(np.float32)
    R = ref_gpu.api()
    for bits in (16, 32):
        pr, pg = str(tmp_path / f"ref{bits}.tif"), str(tmp_path / f"got{bits}.tif")
        R.writetifstack(pr, vol, bits)
        libapi.writetifstack(pg, vol, bits)
        assert R.gettifinfo(pg) == libapi.gettifinfo(pr) == (bits, (20, 12, 5))
        a, b, c, d = R.readtifstack(pr), R.readtifstack(pg), libapi.readtifstack(pr), libapi.readtifstack(pg)
        assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d)
        assert np.array_equal(a, vol if bits == 32 else np.floor(vol))
