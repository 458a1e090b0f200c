"""Slab-decomposed distributed Richardson-Lucy deconvolution: ONE volume over P GPUs (config 3).

One process per GPU (torch.distributed, NCCL).  Decomposition:

  real volumes A, B, E      rank r owns rows y in [r*Y/P, (r+1)*Y/P)      [X][Y/P][Z]
  half spectrum, "slabs"    same rows, all kx                             [X/2+1][Y/P][Z]
  half spectrum, "planes"   rank r owns a contiguous range of kx planes   [np_r][Y][Z]
  OTFs                      plane layout (each rank keeps only its planes)

Per convolution: fused X pencils locally on the slab (csrc/fft_fast.cuh k_xpassP) -> all-to-all ->
Y / Z passes and OTF product locally on whole planes -> all-to-all -> X pencils.  Two all-to-alls
per convolution, eight per dual-view iteration, each moving (P-1)/P of the local spectrum
(4*N*(P-1)/P^2 bytes per GPU).  All arithmetic is done by the same sm_100a kernels as the
single-GPU path through include/milb_capi.h (milb_dslab_*); torch provides device memory, the
stream and the NCCL all-to-all.  The reference has no multi-GPU path at all.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with milb status {rc}")


def plane_counts(nplanes: int, world: int):
    """Contiguous split of the X/2+1 kx-planes over the ranks (first ranks get the extra ones)."""
    base, extra = divmod(nplanes, world)
    return [base + (1 if r < extra else 0) for r in range(world)]


class SlabLayout:
    """Pure index bookkeeping of the two spectrum layouts and of the all-to-all that maps one
    into the other.  Works on any torch device (the gloo tests run it on CPU tensors)."""

    def __init__(self, fft_shape, rank, world):
        self.X, self.Y, self.Z = (int(v) for v in fft_shape)
        if self.Y % world:
            raise ValueError("Y must be divisible by the number of ranks")
        self.rank, self.world = rank, world
        self.ny = self.Y // world
        self.y0 = rank * self.ny
        self.nplanes = self.X // 2 + 1
        self.counts = plane_counts(self.nplanes, world)
        self.np = self.counts[rank]
        self.p0 = sum(self.counts[:rank])
        self.row = self.ny * self.Z * 2          # floats of one (kx, my rows) block

    # element counts (float32) of the blocks exchanged with each peer
    def slab_splits(self):
        return [c * self.row for c in self.counts]

    def plane_splits(self):
        return [self.np * self.row] * self.world

    def to_planes(self, slab, planes, scratch, group=None):
        """slab [nplanes][ny][Z][2] -> planes [np][Y][Z][2] (my planes, all rows)."""
        import torch
        import torch.distributed as dist
        recv = scratch[: self.world * self.np * self.row]
        if self.world == 1:
            recv.copy_(slab.reshape(-1))
        else:
            dist.all_to_all_single(recv, slab.reshape(-1), self.plane_splits(), self.slab_splits(), group=group)
        # received as [source rank][np][ny][Z][2]; rows of source s are y in [s*ny, (s+1)*ny)
        src = recv.view(self.world, self.np, self.ny, self.Z, 2)
        planes.view(self.np, self.world, self.ny, self.Z, 2).copy_(src.permute(1, 0, 2, 3, 4))
        return planes

    def to_slabs(self, planes, slab, scratch, group=None):
        """planes [np][Y][Z][2] -> slab [nplanes][ny][Z][2] (all planes, my rows)."""
        import torch.distributed as dist
        send = scratch[: self.world * self.np * self.row]
        send.view(self.world, self.np, self.ny, self.Z, 2).copy_(planes.view(self.np, self.world, self.ny, self.Z, 2).permute(1, 0, 2, 3, 4))
        if self.world == 1:
            slab.reshape(-1).copy_(send)
        else:
            dist.all_to_all_single(slab.reshape(-1), send, self.slab_splits(), self.plane_splits(), group=group)
        return slab


class DistDecon:
    """Distributed single- or dual-view RL deconvolution of one FFT-box-sized volume."""

    def __init__(self, fft_shape, nviews=1, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.L = SlabLayout(fft_shape, self.rank, self.world)
        L = self.L
        self.nviews = nviews
        self._h = C.c_void_p()
        size = (C.c_uint * 3)(L.Z, L.Y, L.X)
        _check(self.lib.milb_dslab_create(C.byref(self._h), size, L.y0, L.ny, L.np), "milb_dslab_create")
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.A = [torch.empty((L.X, L.ny, L.Z), **f32) for _ in range(nviews)]
        self.E = torch.empty((L.X, L.ny, L.Z), **f32)
        self.slab = torch.empty((L.nplanes, L.ny, L.Z, 2), **f32)
        npmax = max(L.counts)
        self.planes = torch.empty((max(L.np, 1), L.Y, L.Z, 2), **f32)
        self.planes2 = torch.empty((max(L.np, 1), L.Z, L.Y, 2), **f32)
        self.scratch = torch.empty(max(self.world * npmax * L.row, 1), **f32)
        self.otf = [[torch.empty((max(L.np, 1), L.Z, L.Y, 2), **f32) for _ in range(2)] for _ in range(nviews)]
        self.nfft = L.X * L.Y * L.Z

    def close(self):
        if self._h:
            self.lib.milb_dslab_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- thin wrappers over the C-ABI ------------------------------------------------------------
    def _st(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def _xpass(self, mode, vol_io, aux):
        _check(self.lib.milb_dslab_xpass(self._h, mode, self._p(vol_io), self._p(aux), self._p(self.slab), self._st()), "milb_dslab_xpass")

    def _planes(self, otf, scale=1.0):
        if self.L.np:
            _check(self.lib.milb_dslab_planes(self._h, self._p(self.planes), self._p(self.planes2), self._p(otf), C.c_float(scale), self._st()),
                   "milb_dslab_planes")

    def _convolve(self, otf):
        """slab spectrum -> planes -> * otf -> slab spectrum (two all-to-alls)"""
        self.L.to_planes(self.slab, self.planes, self.scratch, self.group)
        self._planes(otf)
        self.L.to_slabs(self.planes, self.slab, self.scratch, self.group)

    # -- public API ------------------------------------------------------------------------------
    def set_psf(self, view, psf, psf_bp=None):
        """OTF and back-projector OTF of one view (genOTFgpu, src/api_subfunc.cu:3270-3307);
        every rank passes the whole (small) PSF."""
        torch = self.torch
        L = self.L
        for which in range(2):
            src = psf_bp if (which == 1 and psf_bp is not None) else psf
            flip = 1 if (which == 1 and psf_bp is None) else 0
            d_psf = torch.as_tensor(np.ascontiguousarray(src, np.float32), device=self.dev)
            size = (C.c_uint * 3)(src.shape[2], src.shape[1], src.shape[0])
            _check(self.lib.milb_dslab_psf_box(self._h, self._p(self.E), self._p(d_psf), size, flip, self._st()), "milb_dslab_psf_box")
            self._xpass(0, self.E, None)
            L.to_planes(self.slab, self.planes, self.scratch, self.group)
            self._planes(None, 1.0 / float(self.nfft))       # forward only; the scaled spectrum is in planes2
            self.otf[view][which].copy_(self.planes2)
            torch.cuda.current_stream().synchronize()         # d_psf is released on return

    def set_image(self, view, slab_img):
        """My rows [X][ny][Z] of the (FFT-box-sized) view: A = max(img, 0.01)."""
        t = self.torch.as_tensor(slab_img, dtype=self.torch.float32, device=self.dev).contiguous()
        assert tuple(t.shape) == tuple(self.A[view].shape)
        _check(self.lib.milb_dslab_elementwise(self._p(self.A[view]), self._p(t), None, t.numel(), 0, self._st()), "clamp")

    def run(self, iterations):
        """decon_singleview_OTF1 / decon_dualview_OTF1 loop (src/api_subfunc.cu:3404-3416, 3634-3660)."""
        nv = self.nviews
        if nv == 1:
            self.E.copy_(self.A[0])
        else:
            _check(self.lib.milb_dslab_elementwise(self._p(self.E), self._p(self.A[0]), self._p(self.A[1]), self.E.numel(), 1, self._st()), "init")
        if iterations == 0:
            return self.E
        self._xpass(0, self.E, None)
        for it in range(iterations):
            for v in range(nv):
                self._convolve(self.otf[v][0])
                self._xpass(1, None, self.A[v])
                self._convolve(self.otf[v][1])
                last = it == iterations - 1 and v == nv - 1
                self._xpass(3 if last else 2, self.E, None)
        return self.E

    def a2a_bytes_per_gpu(self):
        """bytes one GPU sends in one all-to-all"""
        L = self.L
        return 4 * (L.nplanes - L.np) * L.row
