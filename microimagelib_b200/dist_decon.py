"""Slab-decomposed distributed Richardson-Lucy deconvolution: ONE volume over P GPUs (config 3).

One process per GPU (torch.distributed, NCCL).  Decomposition:

  real volumes A, B, E      rank r owns rows y in [r*Y/P, (r+1)*Y/P)      [X][Y/P][Z]
  half spectrum, "slabs"    same rows, all kx                             [X/2+1][Y/P][Z]
  half spectrum, "planes"   rank r owns a contiguous range of kx planes   [np_r][Y][Z]
  OTFs                      plane layout (each rank keeps only its planes)

Per convolution: fused X pencils locally on the slab (csrc/fft_fast.cuh k_xpassP) -> exchange ->
Y / Z passes and OTF product locally on whole planes -> exchange -> X pencils.  Two exchanges per
convolution, eight per dual-view iteration, each moving (P-1)/P of the local spectrum
(4*N*(P-1)/P^2 bytes per GPU).  Two implementations of the exchange:

  fused (default)  the exchange is folded into the kernels: with every rank's slab and plane
                   buffers mapped into every process (CUDA IPC over NVLink), the X pass stores each
                   spectrum row straight into the plane buffer of the rank that owns it and the
                   last plane pass stores each result row straight into its owner's slab buffer --
                   the transfer overlaps the butterflies tile by tile, there is no all-to-all and
                   no re-layout copy; a one-element NCCL all-reduce is the barrier between phases.
  nccl             torch.distributed.all_to_all_single + two re-layout copies (the baseline; also
                   used for the one-time OTF generation).

All arithmetic is done by the same sm_100a kernels as the single-GPU path through
include/milb_capi.h (milb_dslab_*); torch provides device memory, the stream and NCCL.  The
reference has no multi-GPU path at all.
"""
from __future__ import annotations

import ctypes as C
import os

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


class _SharedBuf:
    """cudaMalloc'd float32 buffer that other ranks can map (CUDA IPC); torch sees it through
    __cuda_array_interface__ (torch does not own the memory)."""

    def __init__(self, lib, nfloats):
        self.lib = lib
        self.n = int(nfloats)
        self.ptr = C.c_void_p()
        _check(lib.milb_dev_alloc(C.byref(self.ptr), self.n * 4), "milb_dev_alloc")
        self.__cuda_array_interface__ = {"shape": (self.n,), "typestr": "<f4", "data": (self.ptr.value, False), "version": 2}

    def handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self.lib.milb_ipc_export(self.ptr, buf), "milb_ipc_export")
        return buf.raw

    def free(self):
        if self.ptr:
            self.lib.milb_dev_free(self.ptr)
            self.ptr = C.c_void_p()


class DistDecon:
    """Distributed single- or dual-view RL deconvolution of one FFT-box-sized volume.
    fused=None picks the fused exchange unless MILB_DIST_FUSED=0."""

    def __init__(self, fft_shape, nviews=1, group=None, device=None, fused=None):
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
        self.fused = (os.environ.get("MILB_DIST_FUSED", "1") != "0") if fused is None else bool(fused)
        self._h = C.c_void_p()
        size = (C.c_uint * 3)(L.Z, L.Y, L.X)
        _check(self.lib.milb_dslab_create(C.byref(self._h), size, L.y0, L.ny, L.np), "milb_dslab_create")
        # the fused exchange needs <= 8 ranks, power-of-two slabs and X-pass tiles that stay inside one row (the library knows its
        # tile width); every rank must own at least one plane; anything else takes the all-to-all path
        if self.fused and (min(L.counts) < 1 or not self.lib.milb_dslab_can_fuse(self._h, self.world)):
            self.fused = False
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.A = [torch.empty((L.X, L.ny, L.Z), **f32) for _ in range(nviews)]
        self.E = torch.empty((L.X, L.ny, L.Z), **f32)
        npmax = max(L.counts)
        self._shared, self._opened = [], []
        if self.fused:
            self._setup_peers()
        else:
            self.slab = torch.empty((L.nplanes, L.ny, L.Z, 2), **f32)
            self.planes = torch.empty((max(L.np, 1), L.Y, L.Z, 2), **f32)
        self.planes2 = torch.empty((max(L.np, 1), L.Z, L.Y, 2), **f32)
        self.scratch = torch.empty(max(self.world * npmax * L.row, 1), **f32)
        self.otf = [[torch.empty((max(L.np, 1), L.Z, L.Y, 2), **f32) for _ in range(2)] for _ in range(nviews)]
        self.nfft = L.X * L.Y * L.Z

    def _setup_peers(self):
        """allocate my slab / plane buffers as IPC-shareable memory, map everybody else's, tell the kernels"""
        torch = self.torch
        import torch.distributed as dist
        L = self.L
        slab_b = _SharedBuf(self.lib, L.nplanes * L.ny * L.Z * 2)
        planes_b = _SharedBuf(self.lib, max(L.np, 1) * L.Y * L.Z * 2)
        self._shared = [slab_b, planes_b]
        self.slab = torch.as_tensor(slab_b, device=self.dev).view(L.nplanes, L.ny, L.Z, 2)
        self.planes = torch.as_tensor(planes_b, device=self.dev).view(max(L.np, 1), L.Y, L.Z, 2)
        slab_ptrs = (C.c_void_p * self.world)()
        planes_ptrs = (C.c_void_p * self.world)()
        if self.world > 1:
            mine = (slab_b.handle(), planes_b.handle())
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=self.group)
        for r in range(self.world):
            if r == self.rank:
                slab_ptrs[r], planes_ptrs[r] = slab_b.ptr.value, planes_b.ptr.value
                continue
            for k, arr in ((0, slab_ptrs), (1, planes_ptrs)):
                p = C.c_void_p()
                _check(self.lib.milb_ipc_open(everyone[r][k], C.byref(p)), "milb_ipc_open")
                self._opened.append(p)
                arr[r] = p.value
        counts = (C.c_int * self.world)(*L.counts)
        _check(self.lib.milb_dslab_set_peers(self._h, self.world, self.rank, planes_ptrs, slab_ptrs, counts), "milb_dslab_set_peers")
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self._barrier()

    def _barrier(self):
        """stream-ordered cross-rank barrier: nobody's later kernels start before everybody's earlier ones finished"""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self._flag, group=self.group)

    def close(self):
        if self._h:
            if self.fused and self.world > 1:
                self.torch.cuda.synchronize()
                self._barrier()                       # nobody still writes into a buffer that is about to go away
                self.torch.cuda.synchronize()
            self.lib.milb_dslab_destroy(self._h)
            self._h = C.c_void_p()
            for p in self._opened:
                self.lib.milb_ipc_close(p)
            self._opened = []
            self.slab = self.planes = None
            for b in self._shared:
                b.free()
            self._shared = []

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

    def _xpass_peer(self, mode, vol_io, aux):
        _check(self.lib.milb_dslab_xpass_peer(self._h, mode, self._p(vol_io), self._p(aux), self._p(self.slab), self._st()), "milb_dslab_xpass_peer")

    def _convolve_fused(self, otf):
        """planes (filled by everybody's X pass) -> * otf -> everybody's slabs; the exchange rides on the kernels' stores"""
        self._barrier()                                   # every rank's X pass has landed in my planes
        if self.L.np:
            _check(self.lib.milb_dslab_planes_peer(self._h, self._p(self.planes), self._p(self.planes2), self._p(otf), self._st()),
                   "milb_dslab_planes_peer")
        self._barrier()                                   # every rank's result rows have landed in my slab

    def run(self, iterations):
        """decon_singleview_OTF1 / decon_dualview_OTF1 loop (src/api_subfunc.cu:3404-3416, 3634-3660)."""
        nv = self.nviews
        if nv == 1:
            self.E.copy_(self.A[0])
        else:
            _check(self.lib.milb_dslab_elementwise(self._p(self.E), self._p(self.A[0]), self._p(self.A[1]), self.E.numel(), 1, self._st()), "init")
        if iterations == 0:
            return self.E
        if self.fused:
            self._barrier()                               # the previous call's readers of `planes` are done everywhere
            self._xpass_peer(0, self.E, None)
            for it in range(iterations):
                for v in range(nv):
                    self._convolve_fused(self.otf[v][0])
                    self._xpass_peer(1, None, self.A[v])
                    self._convolve_fused(self.otf[v][1])
                    last = it == iterations - 1 and v == nv - 1
                    self._xpass_peer(3 if last else 2, self.E, None)
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
