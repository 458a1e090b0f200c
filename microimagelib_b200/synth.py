"""Seeded synthetic bead volumes, Gaussian PSFs and known-affine pairs.
Pure numpy/scipy; used by tests and bench.py (never by the product path)."""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft

SEED_A, SEED_B, SEED_NOISE = 20260, 20261, 20262


def gaussian_psf(shape=(65, 65, 65), sigma_zyx=(4.0, 2.0, 2.0)) -> np.ndarray:
    """Gaussian PSF, array axes (slices, H, W); sigma given per array axis."""
    ax = [np.arange(n, dtype=np.float64) - (n // 2) for n in shape]
    g = [np.exp(-0.5 * (a / s) ** 2) for a, s in zip(ax, sigma_zyx)]
    psf = g[0][:, None, None] * g[1][None, :, None] * g[2][None, None, :]
    return (psf / psf.sum()).astype(np.float32)


def beads(shape, seed=SEED_A, density=1.0 / 16384, intensity=2.5e5, margin=8) -> np.ndarray:
    """Point sources on a zero background, float64."""
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    k = max(1, int(round(n * density)))
    vol = np.zeros(shape, np.float64)
    lo = [min(margin, s // 4) for s in shape]
    pos = [rng.integers(l, s - l, size=k) for l, s in zip(lo, shape)]
    np.add.at(vol, tuple(pos), intensity)
    return vol


def blur(vol: np.ndarray, psf: np.ndarray) -> np.ndarray:
    """Circular convolution with the PSF centred at floor(P/2), float64 FFT."""
    box = np.zeros(vol.shape, np.float64)
    idx = []
    for f, p in zip(vol.shape, psf.shape):
        d = np.arange(p) - p // 2
        keep = np.abs(d) < f // 2 if p > f else np.ones(p, bool)
        idx.append((d % f, keep))
    sub = psf.astype(np.float64)[np.ix_(*[k for _, k in idx])]
    box[np.ix_(*[d[k] for d, k in idx])] = sub
    box /= box.sum()
    w = None
    return sfft.irfftn(sfft.rfftn(vol, workers=-1) * sfft.rfftn(box, workers=-1), s=vol.shape, workers=-1)


def bead_image(shape, psf, seed=SEED_A, noise_seed=SEED_NOISE, background=100.0, noise=True, **kw) -> np.ndarray:
    """beads (*) psf + background with Poisson noise, float32."""
    clean = blur(beads(shape, seed, **kw), psf) + background
    clean = np.maximum(clean, 0.0)
    if noise:
        clean = np.random.default_rng(noise_seed).poisson(clean).astype(np.float64)
    return clean.astype(np.float32)


def affine_matrix(rot_z_deg=2.0, scale=(1.02, 0.99, 1.0), shift=(3.5, -2.25, 1.75), center=None) -> np.ndarray:
    """3x4 row-major matrix (x, y, z order = W, H, slices) mapping target voxel -> source voxel:
    rotation about z through `center`, per-axis scale, then shift."""
    a = np.deg2rad(rot_z_deg)
    R = np.array([[np.cos(a), np.sin(a), 0], [-np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float64)
    M = np.diag(scale) @ R
    c = np.zeros(3) if center is None else np.asarray(center, np.float64)
    t = c - M @ c + np.asarray(shift, np.float64)
    return np.concatenate([M, t[:, None]], axis=1).astype(np.float32).reshape(12)


def warp_exact(vol: np.ndarray, m12: np.ndarray, out_shape=None) -> np.ndarray:
    """float64 trilinear resampling, out(x,y,z) = vol(M*(x,y,z,1)), zero outside; chunked over z."""
    from scipy.ndimage import map_coordinates
    out_shape = tuple(out_shape or vol.shape)
    M = np.asarray(m12, np.float64).reshape(3, 4)
    sz, sy, sx = out_shape
    out = np.empty(out_shape, np.float32)
    y, x = np.meshgrid(np.arange(sy, dtype=np.float64), np.arange(sx, dtype=np.float64), indexing="ij")
    for z in range(sz):
        cx = M[0, 0] * x + M[0, 1] * y + M[0, 2] * z + M[0, 3]
        cy = M[1, 0] * x + M[1, 1] * y + M[1, 2] * z + M[1, 3]
        cz = M[2, 0] * x + M[2, 1] * y + M[2, 2] * z + M[2, 3]
        out[z] = map_coordinates(vol, [cz, cy, cx], order=1, mode="constant", cval=0.0)
    return out


def invert_affine(m12) -> np.ndarray:
    """Inverse of a 3x4 affine (implicit last row 0 0 0 1), float64 -> float32."""
    M = np.vstack([np.asarray(m12, np.float64).reshape(3, 4), [0, 0, 0, 1]])
    return np.linalg.inv(M)[:3].astype(np.float32).reshape(12)


def shift_zero_fill(vol, shift):
    """out[z, y, x] = vol[z - dz, y - dy, x - dx] with zeros shifted in; shift = (dx, dy, dz) integers"""
    vol = np.asarray(vol)
    out = np.zeros_like(vol)
    dx, dy, dz = (int(v) for v in shift)

    def rng(n, d):
        lo, hi = max(0, d), min(n, n + d)
        return slice(lo, hi), slice(lo - d, hi - d)
    (zo, zi), (yo, yi), (xo, xi) = rng(vol.shape[0], dz), rng(vol.shape[1], dy), rng(vol.shape[2], dx)
    if zo.start < zo.stop and yo.start < yo.stop and xo.start < xo.stop:
        out[zo, yo, xo] = vol[zi, yi, xi]
    return out
