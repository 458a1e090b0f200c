"""Python mirror of the reference's public C API (include/libapi.h), bound with ctypes.

Same names, argument meaning and return conventions as the C functions; numpy arrays stand in
for the caller-owned host buffers.  Volumes are float32 C-order arrays of shape
``(slices, H, W)`` -- the same bytes as the reference's x-fastest layout with
``imSize = {W, H, slices}``.  Everything runs on the GPU through ``lib/libapi.so``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_F = C.POINTER(C.c_float)
_U = C.POINTER(C.c_uint)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(_F)


def _size(shape):
    """numpy shape (S, H, W) -> reference size triple {W, H, S}"""
    s = (C.c_uint * 3)(int(shape[2]), int(shape[1]), int(shape[0]))
    return s


def snap_transform_size(n: int) -> int:
    return int(_lib.load().milb_snap_transform_size(int(n)))


def decon_singleview(img, psf, itNumForDecon, initialFlag=False, deviceNum=0, gpuMemMode=1, verbose=False,
                     flagUnmatch=False, psf_bp=None, out=None):
    """decon_singleview (include/libapi.h:42-43).  Returns (h_decon, status, deconRecords).
    `out` may be a caller-owned float32 buffer (e.g. pinned memory), as in the C API."""
    lib = _lib.load()
    img, psf = _f32(img), _f32(psf)
    out = np.empty_like(img) if out is None else out
    rec = np.zeros(10, np.float32)
    bp = _f32(psf_bp) if psf_bp is not None else psf
    st = lib.decon_singleview(_fp(out), _fp(img), _size(img.shape), _fp(psf), _size(psf.shape), bool(initialFlag),
                              int(itNumForDecon), int(deviceNum), int(gpuMemMode), bool(verbose), _fp(rec),
                              bool(flagUnmatch), _fp(bp))
    return out, st, rec


def decon_dualview(img1, img2, psf1, psf2, itNumForDecon, initialFlag=False, deviceNum=0, gpuMemMode=1, verbose=False,
                   flagUnmatch=False, psf_bp1=None, psf_bp2=None):
    """decon_dualview (include/libapi.h:45-46).  Returns (h_decon, status, deconRecords)."""
    lib = _lib.load()
    img1, img2, psf1, psf2 = _f32(img1), _f32(img2), _f32(psf1), _f32(psf2)
    out = np.empty_like(img1)
    rec = np.zeros(10, np.float32)
    bp1 = _f32(psf_bp1) if psf_bp1 is not None else psf1
    bp2 = _f32(psf_bp2) if psf_bp2 is not None else psf2
    st = lib.decon_dualview(_fp(out), _fp(img1), _fp(img2), _size(img1.shape), _fp(psf1), _fp(psf2), _size(psf1.shape),
                            bool(initialFlag), int(itNumForDecon), int(deviceNum), int(gpuMemMode), bool(verbose), _fp(rec),
                            bool(flagUnmatch), _fp(bp1), _fp(bp2))
    return out, st, rec


def reg3d(img1, img2, regChoice=2, regMethod=6, inputTmx=False, iTmx=None, FTOL=1e-4, itLimit=3000, deviceNum=0,
          gpuMemMode=1, verbose=False):
    """reg3d (include/libapi.h:35-36).  img1 = target, img2 = source.
    Returns (h_reg, iTmx, status, records)."""
    lib = _lib.load()
    img1, img2 = _f32(img1), _f32(img2)
    out = np.zeros_like(img1)
    tmx = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32) if iTmx is None else _f32(iTmx).reshape(12).copy()
    rec = np.zeros(11, np.float32)
    st = lib.reg3d(_fp(out), _fp(tmx), _fp(img1), _fp(img2), _size(img1.shape), _size(img2.shape), int(regChoice),
                   int(regMethod), bool(inputTmx), float(FTOL), int(itLimit), int(deviceNum), int(gpuMemMode), bool(verbose),
                   _fp(rec))
    return out, tmx, st, rec


def reg2d(img1, img2, regChoice=2, flagTmx=False, iTmx=None, FTOL=1e-4, itLimit=3000, deviceNum=0, gpuMemMode=1, verbose=False):
    """reg2d (include/libapi.h).  Images are (H, W).  Returns (h_reg, iTmx[6], status, records)."""
    lib = _lib.load()
    img1, img2 = _f32(img1), _f32(img2)
    out = np.zeros_like(img1)
    tmx = np.array([1, 0, 0, 0, 1, 0], np.float32) if iTmx is None else _f32(iTmx).reshape(6).copy()
    rec = np.zeros(11, np.float32)
    s1 = (C.c_uint * 3)(img1.shape[1], img1.shape[0], 1)
    s2 = (C.c_uint * 3)(img2.shape[1], img2.shape[0], 1)
    st = lib.reg2d(_fp(out), _fp(tmx), _fp(img1), _fp(img2), s1, s2, int(regChoice), bool(flagTmx), float(FTOL), int(itLimit),
                   int(deviceNum), int(gpuMemMode), bool(verbose), _fp(rec))
    return out, tmx, st, rec


def checkmatrix(iTmx, sx, sy, sz) -> bool:
    m = _f32(iTmx).reshape(12)
    return bool(_lib.load().checkmatrix(_fp(m), int(sx), int(sy), int(sz)))


def atrans3dgpu(img2, iTmx, out_shape=None, deviceNum=0):
    """atrans3dgpu (include/libapi.h:30): warp img2 into a volume of out_shape by iTmx."""
    img2 = _f32(img2)
    out_shape = tuple(out_shape or img2.shape)
    out = np.zeros(out_shape, np.float32)
    m = _f32(iTmx).reshape(12)
    st = _lib.load().atrans3dgpu(_fp(out), _fp(m), _fp(img2), _size(out_shape), _size(img2.shape), int(deviceNum))
    return out, st


def atrans3dgpu_16bit(img2, iTmx, out_shape=None, deviceNum=0):
    img2 = np.ascontiguousarray(img2, dtype=np.uint16)
    out_shape = tuple(out_shape or img2.shape)
    out = np.zeros(out_shape, np.uint16)
    m = _f32(iTmx).reshape(12)
    US = C.POINTER(C.c_ushort)
    st = _lib.load().atrans3dgpu_16bit(out.ctypes.data_as(US), _fp(m), img2.ctypes.data_as(US), _size(out_shape),
                                       _size(img2.shape), int(deviceNum))
    return out, st


def alignsize3d(vol, out_shape, gpuMemMode=1):
    """alignsize3d (include/libapi.h:62): centred crop / zero pad to out_shape (S, H, W)."""
    vol = _f32(vol)
    out = np.zeros(out_shape, np.float32)
    # the C function takes (sx, sy, sz) with sx the slowest axis of its kernel == numpy axis 0
    st = _lib.load().alignsize3d(_fp(out), _fp(vol), out_shape[0], out_shape[1], out_shape[2], vol.shape[0], vol.shape[1],
                                 vol.shape[2], int(gpuMemMode))
    return out, st


def imresize3d(vol, out_shape, deviceNum=0):
    """imresize3d (include/libapi.h:64): resample to out_shape (S, H, W)."""
    vol = _f32(vol)
    out = np.zeros(out_shape, np.float32)
    st = _lib.load().imresize3d(_fp(out), _fp(vol), out_shape[2], out_shape[1], out_shape[0], vol.shape[2], vol.shape[1],
                                vol.shape[0], int(deviceNum))
    return out, st


def imoperation3D(vol, opChoice, deviceNum=0):
    """imoperation3D (include/libapi.h:67): opChoice 1 / 2 = +90 / -90 degrees about Y."""
    vol = _f32(vol)
    size_in = _size(vol.shape)
    size_out = (C.c_uint * 3)(*list(size_in))
    out = np.zeros(vol.size, np.float32)
    st = _lib.load().imoperation3D(_fp(out), size_out, _fp(vol), size_in, int(opChoice), int(deviceNum))
    if opChoice == 0:
        return vol.copy(), st
    return out.reshape(size_out[2], size_out[1], size_out[0]), st


def mp2dgpu(vol, flagZProj=True, flagXProj=True, flagYProj=True):
    """mp2dgpu (include/libapi.h:54).  Returns (zproj (H,W), xproj (S,H), yproj (W,S), status)."""
    vol = _f32(vol)
    sz, sy, sx = vol.shape
    buf = np.zeros(sx * sy + sy * sz + sz * sx, np.float32)
    size_mp = (C.c_uint * 6)()
    st = _lib.load().mp2dgpu(_fp(buf), size_mp, _fp(vol), _size(vol.shape), bool(flagZProj), bool(flagXProj), bool(flagYProj))
    zp = buf[: sx * sy].reshape(sy, sx)
    xp = buf[sx * sy: sx * sy + sy * sz].reshape(sz, sy)
    yp = buf[sx * sy + sy * sz:].reshape(sx, sz)
    return zp, xp, yp, st


def mip3dgpu(vol, rAxis, projectNum):
    """mip3dgpu (include/libapi.h:58).  Returns (stack (projectNum, rows, cols), status)."""
    vol = _f32(vol)
    sz, sy, sx = vol.shape
    if rAxis == 1:
        R = int(round(np.sqrt(float(sy * sy + sz * sz))))
        shape = (projectNum, R, sx)
    else:
        R = int(round(np.sqrt(float(sx * sx + sz * sz))))
        shape = (projectNum, sy, R)
    buf = np.zeros(shape, np.float32)
    size_mp = (C.c_uint * 3)()
    st = _lib.load().mip3dgpu(_fp(buf), size_mp, _fp(vol), _size(vol.shape), int(rAxis), int(projectNum))
    return buf, st


def gettifinfo(path):
    size = (C.c_uint * 3)()
    bits = _lib.load().gettifinfo(str(path).encode(), size)
    return int(bits), (int(size[0]), int(size[1]), int(size[2]))


def readtifstack(path):
    bits, (w, h, s) = gettifinfo(path)
    out = np.zeros((s, h, w), np.float32)
    size = (C.c_uint * 3)()
    _lib.load().readtifstack(_fp(out), str(path).encode(), size)
    return out


def writetifstack(path, vol, bitPerSample=16):
    vol = _f32(vol)
    _lib.load().writetifstack(str(path).encode(), _fp(vol), _size(vol.shape), int(bitPerSample))


def fusion_dualview_status():
    """The reference's fusion_dualview returns 1 without working (src/api_decon.cpp:1133-1136)."""
    lib = _lib.load()
    z = C.cast(None, _F)
    zu = C.cast(None, _U)
    return lib.fusion_dualview(z, z, z, z, z, z, z, zu, zu, z, z, 0, False, 0, 0.0, 0, z, z, zu, 0, 0, 1, False, z, False, z, z)
