"""GPU parity of the "next" rows around the hot path: rotation about Y, centred
crop/pad, resampling, maximum-intensity projections (bit exact against numpy restatements of
include/cukernel.cuh:394-453,754-770 and src/apifunc.cpp:396-644)."""
import numpy as np
import pytest

from microimagelib_b200 import synth

pytestmark = pytest.mark.gpu


def _vol(shape=(13, 21, 34), seed=0):
    return (np.random.default_rng(seed).random(shape) * 1000).astype(np.float32)


def test_rotation_about_y_both_directions():
    from microimagelib_b200 import libapi
    v = _vol()
    sz, sy, sx = v.shape
    # +90: out[x'=k, y'=j, z'=sx-1-i] = in[i,j,k]   (numpy axes are (z, y, x))
    got, st = libapi.imoperation3D(v, 1)
    want = np.transpose(v, (2, 1, 0))[::-1, :, :]         # [z'=sx-1-i][y'][x'=k]
    assert st == 0 and got.shape == (sx, sy, sz) and np.array_equal(got, want)
    # -90: out[x'=sz-1-k, y'=j, z'=i] = in[i,j,k]
    got, st = libapi.imoperation3D(v, 2)
    want = np.transpose(v, (2, 1, 0))[:, :, ::-1]
    assert st == 0 and np.array_equal(got, want)
    back, _ = libapi.imoperation3D(got, 1)
    assert np.array_equal(back, v)
    _, st = libapi.imoperation3D(v, 7)
    assert st == 1


def test_alignsize_matches_oracle():
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    v = _vol((9, 12, 15))
    for out_shape in ((9, 12, 15), (5, 16, 15), (12, 7, 20), (8, 13, 14)):
        got, st = libapi.alignsize3d(v, out_shape)
        assert st == 0 and np.array_equal(got, do.align_size(v, out_shape))
    _, st = libapi.alignsize3d(v, (9, 12, 15), gpuMemMode=5)
    assert st == 1


def test_imresize_is_the_scaled_warp():
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    v = _vol((10, 14, 18))
    out_shape = (25, 14, 18)                # anisotropic z -> isotropic, like spimFusion's pre-processing
    got, st = libapi.imresize3d(v, out_shape)
    m = np.zeros(12, np.float32)
    m[0] = np.float32(18) / np.float32(18)
    m[5] = np.float32(14) / np.float32(14)
    m[10] = np.float32(10) / np.float32(25)
    assert st == 0 and np.array_equal(got, ro.affine_warp(v, m, out_shape))


def test_mp2d_projections_and_flag_quirk():
    from microimagelib_b200 import libapi
    v = _vol() - 300.0                      # negative values: the accumulator starts at 0 (cukernel.cuh:401)
    zp, xp, yp, st = libapi.mp2dgpu(v, True, True, True)
    assert st == 0
    assert np.array_equal(zp, np.maximum(v.max(axis=0), 0))               # [y][x]
    assert np.array_equal(xp, np.maximum(v.max(axis=2), 0))               # [z][y]
    assert np.array_equal(yp, np.maximum(v.max(axis=1), 0).T)             # [x][z]
    # the Y projection is gated by flagZProj, not flagYProj (src/apifunc.cpp:498)
    zp, xp, yp, _ = libapi.mp2dgpu(v, False, True, True)
    assert not zp.any() and not yp.any() and xp.any()
    zp, xp, yp, _ = libapi.mp2dgpu(v, True, False, False)
    assert zp.any() and yp.any() and not xp.any()


def _rot2matrix(theta, sx, sy, sz, axis):
    """numpy float32 restatement of rot2matrix, src/api_subfunc.cu:626-713"""
    from oracle import reg_oracle as ro
    c, s = np.float32(np.cos(np.float32(theta))), np.float32(np.sin(np.float32(theta)))
    t1 = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    t3 = t1.copy()
    if axis == 1:
        t1[7], t1[11] = sy // 2, sz // 2
        r = np.array([1, 0, 0, 0, 0, c, s, 0, 0, -s, c, 0], np.float32)
        n = int(round(np.sqrt(float(sy * sy + sz * sz))))
        t3[7] = t3[11] = int(-n / 2)
    else:
        t1[3], t1[11] = sx // 2, sz // 2
        r = np.array([c, 0, -s, 0, 0, 1, 0, 0, s, 0, c, 0], np.float32)
        n = int(round(np.sqrt(float(sx * sx + sz * sz))))
        t3[3] = t3[11] = int(-n / 2)
    return ro.matrixmultiply(ro.matrixmultiply(t1, r), t3), n


@pytest.mark.parametrize("axis", [1, 2])
def test_mip3d_matches_warp_then_project(axis):
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    v = _vol((11, 14, 17), 4)
    sz, sy, sx = v.shape
    nproj = 5
    got, st = libapi.mip3dgpu(v, axis, nproj)
    assert st == 0
    step = np.float32(3.14159 * 2 / np.float32(nproj))
    for k in range(nproj):
        ang = np.float32(step * np.float32(k))
        m, n = _rot2matrix(ang, sx, sy, sz, axis)
        out_shape = (n, n, sx) if axis == 1 else (n, sy, n)
        rot = ro.affine_warp(v, m, out_shape)
        want = np.maximum(rot.max(axis=0), 0)
        # cos/sin of the angle come from numpy here and from cosf/sinf in the library: allow the few
        # voxels whose 8-bit weight flips; everything else is bit exact
        diff = np.abs(got[k] - want)
        assert (diff > 0).mean() < 0.02 and diff.max() < 20.0


def test_atrans_16bit_is_point_sampled():
    from microimagelib_b200 import libapi
    v = (np.random.default_rng(2).random((6, 8, 10)) * 60000).astype(np.uint16)
    m = np.array([1, 0, 0, 0.4, 0, 1, 0, -0.3, 0, 0, 1, 0.2], np.float32)   # sub-voxel shift
    got, st = libapi.atrans3dgpu_16bit(v, m)
    z, y, x = np.meshgrid(np.arange(6), np.arange(8), np.arange(10), indexing="ij")
    f = np.float32
    tx = (x.astype(f) + f(0.4)) + f(0.5)
    ty = (y.astype(f) + f(-0.3)) + f(0.5)
    tz = (z.astype(f) + f(0.2)) + f(0.5)
    ok = (tx >= 0) & (tx < 10) & (ty >= 0) & (ty < 8) & (tz >= 0) & (tz < 6)
    want = np.where(ok, v[np.clip(np.floor(tz).astype(int), 0, 5), np.clip(np.floor(ty).astype(int), 0, 7), np.clip(np.floor(tx).astype(int), 0, 9)], 0)
    assert st == 0 and np.array_equal(got, want.astype(np.uint16))


def test_libapi_accepts_device_pointers():
    """Extension of this backend: every volume argument of the libapi.h entry points may be a device pointer (used in place,
    written in place).  Same results as with host buffers, bit for bit."""
    import ctypes as C
    import torch
    from microimagelib_b200 import _lib, libapi, synth
    lib = _lib.load()
    F, U = C.POINTER(C.c_float), C.POINTER(C.c_uint)
    rng = np.random.default_rng(3)
    vol = (rng.random((20, 28, 36)) * 4000).astype(np.float32)
    d_vol = torch.from_numpy(vol).cuda()

    def fp(t):
        return C.cast(t.data_ptr(), F)

    size = (C.c_uint * 3)(36, 28, 20)
    # rotation
    want, _ = libapi.imoperation3D(vol, 1)
    d_out = torch.empty(vol.size, dtype=torch.float32, device="cuda")
    so = (C.c_uint * 3)()
    assert lib.imoperation3D(fp(d_out), so, fp(d_vol), size, 1, 0) == 0
    assert np.array_equal(d_out.cpu().numpy().reshape(so[2], so[1], so[0]), want)
    # resampling (the warp)
    want, _ = libapi.imresize3d(vol, (30, 28, 36))
    d_out = torch.empty((30, 28, 36), dtype=torch.float32, device="cuda")
    assert lib.imresize3d(fp(d_out), fp(d_vol), 36, 28, 30, 36, 28, 20, 0) == 0
    assert np.array_equal(d_out.cpu().numpy(), want)
    # 2-D projections: device volume in, host projections out
    zp, xp, yp, _ = libapi.mp2dgpu(vol)
    buf = np.zeros(36 * 28 + 28 * 20 + 20 * 36, np.float32)
    smp = (C.c_uint * 6)()
    assert lib.mp2dgpu(buf.ctypes.data_as(F), smp, fp(d_vol), size, True, True, True) == 0
    assert np.array_equal(buf[:36 * 28].reshape(28, 36), zp) and np.array_equal(buf[36 * 28:36 * 28 + 28 * 20].reshape(20, 28), xp)
    # rotating projections: device volume in, device stack out
    want, _ = libapi.mip3dgpu(vol, 2, 5)
    d_mp = torch.zeros(want.shape, dtype=torch.float32, device="cuda")
    s3 = (C.c_uint * 3)()
    assert lib.mip3dgpu(fp(d_mp), s3, fp(d_vol), size, 2, 5) == 0
    assert np.array_equal(d_mp.cpu().numpy(), want)
    # deconvolution and registration
    psf = synth.gaussian_psf((9, 9, 9), (2, 2, 2))
    img = synth.bead_image((16, 32, 32), psf, density=1 / 256.0, seed=9)
    want, st, _ = libapi.decon_singleview(img, psf, 4)
    d_img, d_dec = torch.from_numpy(img).cuda(), torch.empty(img.shape, dtype=torch.float32, device="cuda")
    rec = np.zeros(10, np.float32)
    isz, psz = (C.c_uint * 3)(32, 32, 16), (C.c_uint * 3)(9, 9, 9)
    assert lib.decon_singleview(fp(d_dec), fp(d_img), isz, psf.ctypes.data_as(F), psz, False, 4, 0, 1, False, rec.ctypes.data_as(F), False,
                                psf.ctypes.data_as(F)) == 0
    assert np.array_equal(d_dec.cpu().numpy(), want)
    m = synth.affine_matrix(rot_z_deg=1.0, scale=(1.01, 0.99, 1.0), shift=(0.5, -0.5, 0.25), center=(16, 16, 8))
    src = synth.warp_exact(img, m)
    reg_w, tmx_w, st, rec_w = libapi.reg3d(img, src, regChoice=2, regMethod=2, FTOL=1e-3, itLimit=300)
    d_src, d_reg = torch.from_numpy(src).cuda(), torch.empty(img.shape, dtype=torch.float32, device="cuda")
    tmx = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    rrec = np.zeros(11, np.float32)
    assert lib.reg3d(fp(d_reg), tmx.ctypes.data_as(F), fp(d_img), fp(d_src), isz, isz, 2, 2, False, 1e-3, 300, 0, 1, False, rrec.ctypes.data_as(F)) == 0
    assert np.array_equal(tmx, tmx_w) and np.array_equal(d_reg.cpu().numpy(), reg_w)
    # 16-bit <-> float conversions: (float)uint16 and the x86 (unsigned short)(int)float truncation, no clamp
    f = np.array([0.0, 0.99, 1.0, 65535.7, 65536.0, 70000.5, -1.0, -0.5, 3e9, -3e9, 123.999], np.float32)
    d_f = torch.from_numpy(f).cuda()
    d_u = torch.zeros(f.size, dtype=torch.int16, device="cuda")
    assert lib.milb_convert_f32_to_u16(d_u.data_ptr(), d_f.data_ptr(), f.size, None) == 0
    torch.cuda.synchronize()
    want16 = np.array([0, 0, 1, 65535, 0, 70000 - 65536, 65535, 0, 0, 0, 123], np.uint16)
    assert np.array_equal(d_u.cpu().numpy().view(np.uint16), want16)
    d_back = torch.zeros(f.size, dtype=torch.float32, device="cuda")
    assert lib.milb_convert_u16_to_f32(d_back.data_ptr(), d_u.data_ptr(), f.size, None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_back.cpu().numpy(), want16.astype(np.float32))
