"""CPU restatement of the Richardson-Lucy path of microImageLib.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py: pinned to the reference's own GPU build -- tests/test_gpu_reference_pinned.py, the committed
tests/golden/reference_vectors.npz -- and by KATs in tests/).

Arrays are numpy, C-order, shape ``(slices, H, W)``: the same bytes as the reference's x-fastest
TIFF layout.  The reference's decon code calls those axes ``(x, y, z)`` with z fastest
(``src/api_decon.cpp:68``), which is exactly numpy's (axis0, axis1, axis2).

All arithmetic is float32 / complex64 like the reference's GPU loop; the FFT is pocketfft
(scipy.fft) in single precision, un-normalised in both directions like cuFFT / FFTW.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.fft as sfft

SMALLVALUE = np.float32(0.01)  # src/api_subfunc.cu:24


def _workers() -> int:
    return int(os.environ.get("MILB_ORACLE_THREADS", os.cpu_count() or 1))


# --------------------------------------------------------------------------- sizes
def snap_transform_size(n: int) -> int:
    """FFT length chosen for an image extent.  Follows src/api_subfunc.cu:57-87."""
    n = int(n)
    n = (n + 15) // 16 * 16
    low = 1 << (n.bit_length() - 1)
    if low == n:
        return n
    hi = low * 2
    if hi <= 128:
        return hi
    return (n + 63) // 64 * 64


def fft_shape_for(im_shape) -> tuple:
    """Per-axis snap of the *image* size only (src/api_decon.cpp:77-79)."""
    return tuple(snap_transform_size(s) for s in im_shape)


# --------------------------------------------------------------------------- pad / crop / flip
def pad_stack(img: np.ndarray, fft_shape) -> np.ndarray:
    """Edge-replicate pad, centred.  src/api_subfunc.cu:1712-1724 + include/cukernel.cuh:699-737."""
    idx = []
    for f, s in zip(fft_shape, img.shape):
        o = (f - s) // 2
        idx.append(np.clip(np.arange(f) - o, 0, s - 1))
    return np.ascontiguousarray(img[np.ix_(*idx)])


def crop_stack(vol: np.ndarray, im_shape) -> np.ndarray:
    """Centred crop.  src/api_subfunc.cu:1735-1747 + include/cukernel.cuh:739-753."""
    sl = []
    for f, s in zip(vol.shape, im_shape):
        o = (f - s) // 2
        sl.append(slice(o, o + s))
    return np.ascontiguousarray(vol[tuple(sl)])


def flip3(psf: np.ndarray) -> np.ndarray:
    """Reverse all three axes.  include/cukernel.cuh:667-677."""
    return np.ascontiguousarray(psf[::-1, ::-1, ::-1])


def _cdiv2(a: int) -> int:
    """C integer division by two (truncation toward zero), for possibly negative a."""
    return int(a / 2) if a < 0 else a // 2


def align_size(vol: np.ndarray, out_shape) -> np.ndarray:
    """Centred crop-or-zero-pad.  src/api_subfunc.cu:1778-1790 + include/cukernel.cuh:754-770.
    out[d] = in[d - (s_out - s_in)/2] or 0 outside (C truncating division)."""
    out = np.zeros(out_shape, dtype=vol.dtype)
    src, dst = [], []
    for so, si in zip(out_shape, vol.shape):
        off = _cdiv2(so - si)
        d = np.arange(so)
        x = d - off
        ok = (x >= 0) & (x < si)
        dst.append(d[ok])
        src.append(x[ok])
    out[np.ix_(*dst)] = vol[np.ix_(*src)]
    return out


def pad_psf(psf: np.ndarray, fft_shape) -> np.ndarray:
    """Scatter the PSF into a zeroed FFT box, centre floor(P/2) -> origin with wrap.
    src/api_subfunc.cu:1690-1701 + include/cukernel.cuh:679-697 (GPU twin: output zero-filled)."""
    out = np.zeros(fft_shape, dtype=psf.dtype)
    idx = []
    for f, p in zip(fft_shape, psf.shape):
        d = np.arange(p) - p // 2
        d[d < 0] += f
        assert ((d >= 0) & (d < f)).all()
        idx.append(d)
    out[np.ix_(*idx)] = psf
    return out


# --------------------------------------------------------------------------- OTF
def rfftn_u(x: np.ndarray) -> np.ndarray:
    """Un-normalised R2C over all three axes, last axis halved (cuFFT R2C semantics)."""
    return sfft.rfftn(x.astype(np.float32, copy=False), workers=_workers()).astype(np.complex64, copy=False)


def irfftn_u(s: np.ndarray, shape) -> np.ndarray:
    """Un-normalised C2R (no 1/N), cuFFT C2R semantics."""
    return sfft.irfftn(s, s=shape, norm="forward", workers=_workers()).astype(np.float32, copy=False)


def gen_otf(psf: np.ndarray, fft_shape) -> np.ndarray:
    """OTF of a PSF.  Follows genOTFgpu, src/api_subfunc.cu:3270-3307.
    PSF / sum (double sum, float reciprocal), fit to box, shift centre to origin, R2C."""
    psf = np.asarray(psf, dtype=np.float32)
    s = float(np.sum(psf, dtype=np.float64))
    psf_n = psf * np.float32(1.0 / s)
    if any(f < p for f, p in zip(fft_shape, psf.shape)):
        boxed = align_size(psf_n, fft_shape)
        padded = pad_psf(boxed, fft_shape)
    else:
        padded = pad_psf(psf_n, fft_shape)
    return rfftn_u(padded)


def gen_otf_pair(psf, fft_shape, unmatch=False, psf_bp=None):
    """(OTF, OTF_bp) as decon_singleview builds them, src/api_decon.cpp:213-223."""
    otf = gen_otf(psf, fft_shape)
    if unmatch:
        otf_bp = gen_otf(psf_bp, fft_shape)
    else:
        otf_bp = gen_otf(flip3(np.asarray(psf, dtype=np.float32)), fft_shape)
    return otf, otf_bp


# --------------------------------------------------------------------------- RL loops
def _rl_half_step(E, A, otf, otf_bp):
    shape = E.shape
    T = irfftn_u(rfftn_u(E) * otf, shape)          # = N * (E conv h)
    T = A / T                                       # no zero guard (cukernel.cuh:194-206)
    T = irfftn_u(rfftn_u(T) * otf_bp, shape)       # N cancels
    E = E * T
    return np.maximum(E, SMALLVALUE)


def rl_single(A: np.ndarray, otf, otf_bp, iters: int, const_init: bool = False) -> np.ndarray:
    """decon_singleview_OTF1, src/api_subfunc.cu:3361-3430."""
    A = np.maximum(np.asarray(A, dtype=np.float32), SMALLVALUE)
    if const_init:
        mean_value = np.float32(np.sum(A, dtype=np.float64))  # the SUM, not the mean (sic) :3382
        E = np.full_like(A, mean_value)
    else:
        E = A.copy()
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for _ in range(iters):
            E = _rl_half_step(E, A, otf, otf_bp)
    return E


def rl_dual(A, B, otf1, otf_bp1, otf2, otf_bp2, iters: int, const_init: bool = False) -> np.ndarray:
    """decon_dualview_OTF1, src/api_subfunc.cu:3587-3674."""
    A = np.maximum(np.asarray(A, dtype=np.float32), SMALLVALUE)
    B = np.maximum(np.asarray(B, dtype=np.float32), SMALLVALUE)
    if const_init:
        s1 = np.float32(np.sum(A, dtype=np.float64))
        s2 = np.float32(np.sum(B, dtype=np.float64))
        E = np.full_like(A, np.float32((s1 + s2) / np.float32(2)))
    else:
        E = (A + B) * np.float32(0.5)               # add3D then multivalue 0.5, :3616-3617
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for _ in range(iters):
            E = _rl_half_step(E, A, otf1, otf_bp1)
            E = _rl_half_step(E, B, otf2, otf_bp2)
    return E


# --------------------------------------------------------------------------- API level
def decon_singleview(img, psf, iters, const_init=False, unmatch=False, psf_bp=None):
    """decon_singleview, src/api_decon.cpp:53-331 (gpuMemMode 1 branch)."""
    img = np.asarray(img, dtype=np.float32)
    fshape = fft_shape_for(img.shape)
    otf, otf_bp = gen_otf_pair(psf, fshape, unmatch, psf_bp)
    need_pad = any(s < f for s, f in zip(img.shape, fshape))
    A = pad_stack(img, fshape) if need_pad else img
    E = rl_single(A, otf, otf_bp, iters, const_init)
    return crop_stack(E, img.shape) if need_pad else E


def decon_dualview(img1, img2, psf1, psf2, iters, const_init=False, unmatch=False,
                   psf_bp1=None, psf_bp2=None):
    """decon_dualview, src/api_decon.cpp:333-704 (gpuMemMode 1 branch)."""
    img1 = np.asarray(img1, dtype=np.float32)
    img2 = np.asarray(img2, dtype=np.float32)
    fshape = fft_shape_for(img1.shape)
    otf1, otf_bp1 = gen_otf_pair(psf1, fshape, unmatch, psf_bp1)
    otf2, otf_bp2 = gen_otf_pair(psf2, fshape, unmatch, psf_bp2)
    need_pad = any(s < f for s, f in zip(img1.shape, fshape))
    A = pad_stack(img1, fshape) if need_pad else img1
    B = pad_stack(img2, fshape) if need_pad else img2
    E = rl_dual(A, B, otf1, otf_bp1, otf2, otf_bp2, iters, const_init)
    return crop_stack(E, img1.shape) if need_pad else E
