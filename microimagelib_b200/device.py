"""Device-resident handles over the milb_* C-ABI (include/milb_capi.h).

Inputs may be numpy arrays (host memory, copied by the library) or torch CUDA tensors (device
pointers are passed through untouched; torch only provides memory and streams).  Volumes are
float32 ``(slices, H, W)``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

_F = C.POINTER(C.c_float)
_D = C.POINTER(C.c_double)


class MilbError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise MilbError(f"{what} failed with milb status {rc}")


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _ptr(a):
    """(void* pointer, on_device flag, keepalive) for a numpy array or a torch tensor"""
    if _is_torch(a):
        assert a.dtype.is_floating_point and a.element_size() == 4 and a.is_contiguous()
        return C.c_void_p(a.data_ptr()), (1 if a.is_cuda else 0), a
    a = np.ascontiguousarray(a, dtype=np.float32)
    return C.c_void_p(a.ctypes.data), 0, a


def _size(shape):
    return (C.c_uint * 3)(int(shape[2]), int(shape[1]), int(shape[0]))


def _stream(stream):
    if stream is None:
        return C.c_void_p(0)
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


def launch_count() -> int:
    return int(_lib.load().milb_launch_count())


class Decon:
    """Richardson-Lucy state for one image size: OTFs, views, estimate (milb_decon_t)."""

    def __init__(self, im_shape, nviews=1, row_conv=None):
        """row_conv: None = the library's default (the in-place row convolution k_zrow where the Z length has a two-stage
        plan); False = the transposing plane kernels (k_ypassT / k_zconvT)."""
        self.lib = _lib.load()
        self.im_shape = tuple(int(s) for s in im_shape)
        self.nviews = nviews
        self._h = C.c_void_p()
        old = os.environ.get("MILB_ZROW")
        if row_conv is not None:
            os.environ["MILB_ZROW"] = "1" if row_conv else "0"
        try:
            _check(self.lib.milb_decon_create(C.byref(self._h), nviews, _size(self.im_shape)), "milb_decon_create")
        finally:
            if row_conv is not None:
                if old is None:
                    os.environ.pop("MILB_ZROW", None)
                else:
                    os.environ["MILB_ZROW"] = old
        fs = (C.c_uint * 3)()
        self.lib.milb_decon_fft_size(self._h, fs)
        self.fft_shape = (int(fs[2]), int(fs[1]), int(fs[0]))

    def close(self):
        if self._h:
            self.lib.milb_decon_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_psf(self, view, psf, psf_bp=None, stream=None):
        p, dev, keep = _ptr(psf)
        if psf_bp is not None:
            pb, devb, keepb = _ptr(psf_bp)
            assert devb == dev
        else:
            pb, keepb = C.c_void_p(0), None
        shape = psf.shape
        _check(self.lib.milb_decon_set_psf(self._h, view, p, pb, _size(shape), 1 if psf_bp is not None else 0, dev, _stream(stream)),
               "milb_decon_set_psf")

    def set_image(self, view, img, stream=None):
        p, dev, keep = _ptr(img)
        assert tuple(img.shape) == self.im_shape
        _check(self.lib.milb_decon_set_image(self._h, view, p, dev, _stream(stream)), "milb_decon_set_image")

    def run(self, iterations, const_init=False, stream=None):
        _check(self.lib.milb_decon_run(self._h, int(iterations), 1 if const_init else 0, _stream(stream)), "milb_decon_run")

    def run_cufft_yardstick(self, iterations, const_init=False, stream=None):
        ms = C.c_float(0)
        _check(self.lib.milb_decon_run_cufft_yardstick(self._h, int(iterations), 1 if const_init else 0, _stream(stream), C.byref(ms)),
               "milb_decon_run_cufft_yardstick")
        return float(ms.value)

    def time_kernels(self, reps=5, stream=None):
        """average ms per launch of (Y-forward, Z-conv, Y-inverse, X ratio, X update), CUDA events around every launch;
        with the fused plane stage slot 0 is the whole stage and slots 1, 2 are zero"""
        ms = np.zeros(5, np.float32)
        _check(self.lib.milb_decon_time_kernels(self._h, int(reps), ms.ctypes.data_as(_F), _stream(stream)), "milb_decon_time_kernels")
        return ms

    def plane_stage_fused(self):
        """True if the plane stage of a convolution runs as one persistent launch (k_planes_fused)"""
        return bool(self.lib.milb_decon_plane_stage_fused(self._h))

    def time_pipe(self, reps=5, stream=None):
        """plane pipeline only: ms of (Y forward, row convolution, Y inverse, whole stage), each on its own stream"""
        ms = np.zeros(4, np.float32)
        _check(self.lib.milb_decon_time_pipe(self._h, int(reps), ms.ctypes.data_as(_F), _stream(stream)), "milb_decon_time_pipe")
        return ms

    def row_convolution(self):
        """True if the Z convolution runs in place along the contiguous axis (k_zrow) instead of on transposed planes"""
        return bool(self.lib.milb_decon_row_convolution(self._h))

    def set_chunk_planes(self, planes):
        _check(self.lib.milb_decon_set_chunk_planes(self._h, int(planes)), "milb_decon_set_chunk_planes")

    def result(self, out=None, stream=None):
        if out is None:
            out = np.empty(self.im_shape, np.float32)
        p, dev, keep = _ptr(out)
        _check(self.lib.milb_decon_get_result(self._h, p, dev, _stream(stream)), "milb_decon_get_result")
        return out


class Reg:
    """Mean-removed target/source pair for ZNCC cost evaluations (milb_reg_t)."""

    def __init__(self, shape, fetch=None):
        """fetch: None = library default (hardware texture unit), "hw" or "sw" (software restatement, bit-identical
        to the CPU oracle)."""
        self.lib = _lib.load()
        self.shape = tuple(int(s) for s in shape)
        self._h = C.c_void_p()
        _check(self.lib.milb_reg_create(C.byref(self._h), _size(self.shape)), "milb_reg_create")
        self.sd_t = None
        if fetch is not None:
            _check(self.lib.milb_reg_set_fetch(self._h, 1 if fetch == "hw" else 0), "milb_reg_set_fetch")

    def close(self):
        if self._h:
            self.lib.milb_reg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_images(self, target, source, stream=None):
        pt, dt, k1 = _ptr(target)
        ps, ds, k2 = _ptr(source)
        assert dt == ds
        _check(self.lib.milb_reg_set_images(self._h, pt, ps, dt, _stream(stream)), "milb_reg_set_images")

    def prepare(self, pre_tmx=None, stream=None):
        sd = C.c_float(0)
        if pre_tmx is not None:
            m = np.ascontiguousarray(pre_tmx, np.float32).reshape(12)
            mp = m.ctypes.data_as(_F)
        else:
            mp = C.cast(None, _F)
        _check(self.lib.milb_reg_prepare(self._h, mp, C.byref(sd), _stream(stream)), "milb_reg_prepare")
        self.sd_t = np.float32(sd.value)
        return self.sd_t

    def cost(self, matrices, stream=None):
        m = np.ascontiguousarray(matrices, np.float32).reshape(-1, 12)
        out = np.zeros(m.shape[0], np.float32)
        _check(self.lib.milb_reg_cost(self._h, m.ctypes.data_as(_F), m.shape[0], out.ctypes.data_as(_F), _stream(stream)),
               "milb_reg_cost")
        return out

    def cost_sums(self, matrices, stream=None):
        m = np.ascontiguousarray(matrices, np.float32).reshape(-1, 12)
        ss = np.zeros(m.shape[0], np.float64)
        st = np.zeros(m.shape[0], np.float64)
        _check(self.lib.milb_reg_cost_sums(self._h, m.ctypes.data_as(_F), m.shape[0], ss.ctypes.data_as(_D), st.ctypes.data_as(_D),
                                           _stream(stream)), "milb_reg_cost_sums")
        return ss, st

    def warp_source(self, tmx, out=None, stream=None):
        m = np.ascontiguousarray(tmx, np.float32).reshape(12)
        if out is None:
            out = np.empty(self.shape, np.float32)
        p, dev, keep = _ptr(out)
        _check(self.lib.milb_reg_warp_source(self._h, m.ctypes.data_as(_F), p, dev, _stream(stream)), "milb_reg_warp_source")
        return out


def affine_warp(src, tmx, out_shape=None, stream=None):
    lib = _lib.load()
    out_shape = tuple(out_shape or src.shape)
    m = np.ascontiguousarray(tmx, np.float32).reshape(12)
    ps, dev, keep = _ptr(src)
    if dev:
        import torch
        out = torch.empty(out_shape, dtype=torch.float32, device=src.device)
    else:
        out = np.empty(out_shape, np.float32)
    po, _, _ = _ptr(out)
    _check(lib.milb_affine_warp(po, _size(out_shape), ps, _size(src.shape), m.ctypes.data_as(_F), dev, _stream(stream)),
           "milb_affine_warp")
    return out


def reg3d_affine(target, source, aff_method, flag_tmx=False, itmx=None, ftol=1e-4, it_limit=3000, verbose=False, stream=None):
    """milb_reg3d_affine on host arrays or CUDA tensors.  Returns (reg, tmx, records)."""
    lib = _lib.load()
    pt, dt, k1 = _ptr(target)
    ps, ds, k2 = _ptr(source)
    assert dt == ds
    if dt:
        import torch
        reg = torch.empty_like(target)
    else:
        reg = np.empty(tuple(target.shape), np.float32)
    pr, _, _ = _ptr(reg)
    tmx = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32) if itmx is None else np.array(itmx, np.float32).reshape(12)
    rec = np.zeros(11, np.float32)
    rc = lib.milb_reg3d_affine(pr, tmx.ctypes.data_as(_F), pt, ps, _size(target.shape), int(aff_method), 1 if flag_tmx else 0,
                               float(ftol), int(it_limit), dt, 1 if verbose else 0, rec.ctypes.data_as(_F), _stream(stream))
    _check(rc, "milb_reg3d_affine")
    return reg, tmx, rec


# ---- pre-alignment (reg3d regChoice 1 / 3 / 4, reg2d) ------------------------------------------
def _size2(shape):
    return (C.c_uint * 2)(int(shape[1]), int(shape[0]))


def _to_device(a):
    import torch
    return a if _is_torch(a) else torch.as_tensor(np.ascontiguousarray(a, np.float32), device="cuda")


def phasor(img1, img2, stream=None):
    """milb_phasor: integer (x, y, z) shift of img2 against img1; (S, H, W) volumes or (H, W) images."""
    lib = _lib.load()
    a, b = _to_device(img1), _to_device(img2)
    shape = tuple(a.shape) if a.dim() == 3 else (1,) + tuple(a.shape)
    sh = (C.c_longlong * 3)()
    _check(lib.milb_phasor(sh, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), _size(shape), _stream(stream)), "milb_phasor")
    return [int(sh[0]), int(sh[1]), int(sh[2])]


def imshift(vol, shift, stream=None):
    """milb_imshift: out[x] = vol[x - shift] with zero fill, shift = (dx, dy, dz); returns a numpy array."""
    import torch
    lib = _lib.load()
    a = _to_device(vol)
    out = torch.empty_like(a)
    sh = (C.c_longlong * 3)(*[int(v) for v in shift])
    _check(lib.milb_imshift(C.c_void_p(out.data_ptr()), C.c_void_p(a.data_ptr()), _size(a.shape), sh, _stream(stream)), "milb_imshift")
    torch.cuda.synchronize()
    return out.cpu().numpy()


class Reg2D:
    """milb_reg2d_*: mean-removed 2-D image pair; matrices are 6 floats (2x3, target -> source)."""

    def __init__(self, img1, img2, stream=None):
        self.lib = _lib.load()
        self._h = C.c_void_p()
        p1, d1, self._k1 = _ptr(img1)
        p2, d2, self._k2 = _ptr(img2)
        assert d1 == d2
        self.shape = tuple(img1.shape)
        sd = C.c_float(0)
        _check(self.lib.milb_reg2d_create(C.byref(self._h), p1, _size2(img1.shape), p2, _size2(img2.shape), d1, C.byref(sd), _stream(stream)),
               "milb_reg2d_create")
        self.sd_t = np.float32(sd.value)

    def close(self):
        if self._h:
            self.lib.milb_reg2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cost(self, matrices, stream=None):
        m = np.ascontiguousarray(matrices, np.float32).reshape(-1, 6)
        out = np.empty(m.shape[0], np.float32)
        _check(self.lib.milb_reg2d_cost(self._h, m.ctypes.data_as(_F), m.shape[0], out.ctypes.data_as(_F), _stream(stream)), "milb_reg2d_cost")
        return out

    def warp(self, tmx, raw_source=True, stream=None):
        m = np.ascontiguousarray(tmx, np.float32).reshape(6)
        out = np.empty(self.shape, np.float32)
        _check(self.lib.milb_reg2d_warp(self._h, m.ctypes.data_as(_F), 1 if raw_source else 0, C.c_void_p(out.ctypes.data), 0, _stream(stream)),
               "milb_reg2d_warp")
        return out


def reg2d_shiftalign(img1, img2, flag_tmx=False, itmx=None, search_y=True, shift_region=0.3, total_step=30.0, stream=None):
    """milb_reg2d_shiftalign on host images.  Returns (reg, tmx, records)."""
    lib = _lib.load()
    p1, d1, k1 = _ptr(img1)
    p2, d2, k2 = _ptr(img2)
    reg = np.empty(tuple(img1.shape), np.float32)
    tmx = np.array([1, 0, 0, 0, 1, 0], np.float32) if itmx is None else np.array(itmx, np.float32).reshape(6)
    rec = np.zeros(9, np.float32)
    assert d1 == 0 and d2 == 0
    _check(lib.milb_reg2d_shiftalign(C.c_void_p(reg.ctypes.data), tmx.ctypes.data_as(_F), p1, _size2(img1.shape), p2, _size2(img2.shape),
                                     1 if flag_tmx else 0, 1 if search_y else 0, float(shift_region), float(total_step), 0,
                                     rec.ctypes.data_as(_F), _stream(stream)), "milb_reg2d_shiftalign")
    return reg, tmx, rec


def reg2d_affine(img1, img2, aff_method=1, flag_tmx=False, itmx=None, ftol=1e-4, it_limit=3000, stream=None):
    """milb_reg2d_affine on host images.  Returns (reg, tmx, records)."""
    lib = _lib.load()
    p1, d1, k1 = _ptr(img1)
    p2, d2, k2 = _ptr(img2)
    assert d1 == 0 and d2 == 0
    reg = np.empty(tuple(img1.shape), np.float32)
    tmx = np.array([1, 0, 0, 0, 1, 0], np.float32) if itmx is None else np.array(itmx, np.float32).reshape(6)
    rec = np.zeros(11, np.float32)
    _check(lib.milb_reg2d_affine(C.c_void_p(reg.ctypes.data), tmx.ctypes.data_as(_F), p1, _size2(img1.shape), p2, _size2(img2.shape),
                                 int(aff_method), 1 if flag_tmx else 0, float(ftol), int(it_limit), 0, rec.ctypes.data_as(_F),
                                 _stream(stream)), "milb_reg2d_affine")
    return reg, tmx, rec


def tex2d_samples(img, coords, hardware=False):
    """test utility: tex2D of a (H, W) image at (n, 2) texture coordinates, hardware unit or software restatement"""
    lib = _lib.load()
    img = np.ascontiguousarray(img, np.float32)
    c = np.ascontiguousarray(coords, np.float32).reshape(-1, 2)
    out = np.empty(c.shape[0], np.float32)
    fn = lib.milb_debug_tex2d_sample
    fn.restype = C.c_int
    fn.argtypes = [_F, _F, C.POINTER(C.c_uint), _F, C.c_int, C.c_int]
    _check(fn(out.ctypes.data_as(_F), img.ctypes.data_as(_F), _size2(img.shape), c.ctypes.data_as(_F), c.shape[0], 1 if hardware else 0),
           "milb_debug_tex2d_sample")
    return out


# ---- host-side helpers exported for parity tests
def p2matrix(x):
    x = np.ascontiguousarray(x, np.float32)
    m = np.zeros(12, np.float32)
    _lib.load().milb_p2matrix(m.ctypes.data_as(_F), x.ctypes.data_as(_F))
    return m


def matrix2p(m):
    m = np.ascontiguousarray(m, np.float32)
    x = np.zeros(13, np.float32)
    _lib.load().milb_matrix2p(m.ctypes.data_as(_F), x.ctypes.data_as(_F))
    return x


def matrixmultiply(m1, m2):
    m1 = np.ascontiguousarray(m1, np.float32)
    m2 = np.ascontiguousarray(m2, np.float32)
    m = np.zeros(12, np.float32)
    _lib.load().milb_matrixmultiply(m.ctypes.data_as(_F), m1.ctypes.data_as(_F), m2.ctypes.data_as(_F))
    return m


def dof9tomatrix(q, dof):
    q = np.ascontiguousarray(q, np.float32)
    m = np.zeros(12, np.float32)
    _lib.load().milb_dof9tomatrix(m.ctypes.data_as(_F), q.ctypes.data_as(_F), int(dof))
    return m


def powell(p, xi, n, ftol, func, it_limit=3000, counter=None):
    """milb_powell on a Python cost function (host only; used by the optimiser parity tests).
    p: 1-indexed float32 array (len n+1), xi: (n, n) float32 array, updated in place.
    `func(x)` gets a 1-indexed numpy copy; `counter` is a ctypes c_int the function may bump.
    Returns (iter, fret)."""
    lib = _lib.load()
    p = np.ascontiguousarray(p, np.float32)
    xi_c = np.ascontiguousarray(xi, np.float32)
    counter = counter if counter is not None else C.c_int(0)
    it = C.c_int(0)
    fret = C.c_float(0)

    def _cb(ptr, user):
        x = np.ctypeslib.as_array(ptr, shape=(n + 1,)).copy()
        return float(func(x))

    cb = _lib.COSTFN(_cb)
    rc = lib.milb_powell(p.ctypes.data_as(_F), xi_c.ctypes.data_as(_F), n, C.c_float(ftol), C.byref(it), C.byref(fret), cb, None,
                         C.byref(counter), int(it_limit))
    _check(rc, "milb_powell")
    xi[...] = xi_c
    return it.value, np.float32(fret.value), p
