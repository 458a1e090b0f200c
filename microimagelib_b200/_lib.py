"""Loads the in-tree CUDA shared library and declares the C prototypes (ctypes).

There is no CPU fallback: if ``lib/libapi.so`` is missing this raises, and every compute entry
point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MILB_LIBAPI: another build of the same library (A/B measurements of compile-time variants); default: the in-tree build
LIB_PATH = os.environ.get("MILB_LIBAPI") or os.path.join(HERE, "lib", "libapi.so")

_F = C.POINTER(C.c_float)
_D = C.POINTER(C.c_double)
_U = C.POINTER(C.c_uint)
_US = C.POINTER(C.c_ushort)
_I = C.POINTER(C.c_int)
_LL = C.c_longlong
_VP = C.c_void_p
COSTFN = C.CFUNCTYPE(C.c_float, _F, _VP)

# name -> (restype, argtypes): exactly the declarations of include/libapi.h and include/milb_capi.h
LIBAPI_PROTOS = {
    "concat": (C.c_char_p, None),  # variadic
    "fexists": (C.c_bool, [C.c_char_p]),
    "gettifinfo": (C.c_ushort, [C.c_char_p, _U]),
    "readtifstack": (None, [_F, C.c_char_p, _U]),
    "writetifstack": (None, [C.c_char_p, _F, _U, C.c_ushort]),
    "readtifstack_16to16": (None, [_US, C.c_char_p, _U]),
    "writetifstack_16to16": (None, [C.c_char_p, _US, _U]),
    "queryDevice": (None, []),
    "reg2d": (C.c_int, [_F, _F, _F, _F, _U, _U, C.c_int, C.c_bool, C.c_float, C.c_int, C.c_int, C.c_int, C.c_bool, _F]),
    "checkmatrix": (C.c_bool, [_F, _LL, _LL, _LL]),
    "atrans3dgpu": (C.c_int, [_F, _F, _F, _U, _U, C.c_int]),
    "atrans3dgpu_16bit": (C.c_int, [_US, _F, _US, _U, _U, C.c_int]),
    "reg3d": (C.c_int, [_F, _F, _F, _F, _U, _U, C.c_int, C.c_int, C.c_bool, C.c_float, C.c_int, C.c_int, C.c_int, C.c_bool, _F]),
    "reg_3dgpu": (C.c_int, [_F, _F, _F, _F, _U, _U, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _F]),
    "decon_singleview": (C.c_int, [_F, _F, _U, _F, _U, C.c_bool, C.c_int, C.c_int, C.c_int, C.c_bool, _F, C.c_bool, _F]),
    "decon_dualview": (C.c_int, [_F, _F, _F, _U, _F, _F, _U, C.c_bool, C.c_int, C.c_int, C.c_int, C.c_bool, _F, C.c_bool, _F, _F]),
    "fusion_dualview": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _U, _U, _F, _F, C.c_int, C.c_bool, C.c_int, C.c_float, C.c_int,
                                  _F, _F, _U, C.c_int, C.c_int, C.c_int, C.c_bool, _F, C.c_bool, _F, _F]),
    "mp2dgpu": (C.c_int, [_F, _U, _F, _U, C.c_bool, C.c_bool, C.c_bool]),
    "mp3dgpu": (C.c_int, [_F, _U, _F, _U, C.c_bool, C.c_bool, C.c_int]),
    "mip3dgpu": (C.c_int, [_F, _U, _F, _U, C.c_int, _LL]),
    "alignsize3d": (C.c_int, [_F, _F, _LL, _LL, _LL, _LL, _LL, _LL, C.c_int]),
    "imresize3d": (C.c_int, [_F, _F, _LL, _LL, _LL, _LL, _LL, _LL, C.c_int]),
    "imoperation3D": (C.c_int, [_F, _U, _F, _U, C.c_int, C.c_int]),
}

CAPI_PROTOS = {
    "milb_version": (C.c_char_p, []),
    "milb_snap_transform_size": (C.c_int, [C.c_int]),
    "milb_launch_count": (_LL, []),
    "milb_decon_create": (C.c_int, [C.POINTER(_VP), C.c_int, _U]),
    "milb_decon_destroy": (None, [_VP]),
    "milb_decon_fft_size": (C.c_int, [_VP, _U]),
    "milb_decon_set_psf": (C.c_int, [_VP, C.c_int, _VP, _VP, _U, C.c_int, C.c_int, _VP]),
    "milb_decon_set_image": (C.c_int, [_VP, C.c_int, _VP, C.c_int, _VP]),
    "milb_decon_run": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "milb_decon_get_result": (C.c_int, [_VP, _VP, C.c_int, _VP]),
    "milb_decon_set_chunk_planes": (C.c_int, [_VP, C.c_int]),
    "milb_decon_plane_stage_fused": (C.c_int, [_VP]),
    "milb_decon_row_convolution": (C.c_int, [_VP]),
    "milb_decon_time_pipe": (C.c_int, [_VP, C.c_int, _F, _VP]),
    "milb_decon_run_host": (C.c_int, [_VP, C.POINTER(_VP), _VP, C.c_int, C.c_int, _VP]),
    "milb_decon_phase_correlate": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "milb_decon_alive": (C.c_int, [_VP]),
    "milb_decon_abandon": (None, [_VP]),
    "milb_decon_psf_matches": (C.c_int, [_VP, C.c_int, _VP, _VP, _U, C.c_int]),
    "milb_decon_cache_release": (None, []),
    "milb_decon_run_cufft_yardstick": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _F]),
    "milb_decon_time_kernels": (C.c_int, [_VP, C.c_int, _F, _VP]),
    "milb_dslab_create": (C.c_int, [C.POINTER(_VP), _U, C.c_int, C.c_int, C.c_int]),
    "milb_dslab_destroy": (None, [_VP]),
    "milb_dslab_xpass": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, _VP]),
    "milb_dslab_planes": (C.c_int, [_VP, _VP, _VP, _VP, C.c_float, _VP]),
    "milb_dslab_psf_box": (C.c_int, [_VP, _VP, _VP, _U, C.c_int, _VP]),
    "milb_dslab_elementwise": (C.c_int, [_VP, _VP, _VP, _LL, C.c_int, _VP]),
    "milb_dslab_can_fuse": (C.c_int, [_VP, C.c_int]),
    "milb_dslab_set_peers": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(_VP), C.POINTER(_VP), _I]),
    "milb_dslab_xpass_peer": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, _VP]),
    "milb_dslab_planes_peer": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "milb_convert_u16_to_f32": (C.c_int, [_VP, _VP, _LL, _VP]),
    "milb_convert_f32_to_u16": (C.c_int, [_VP, _VP, _LL, _VP]),
    "milb_host_alloc": (C.c_int, [C.POINTER(_VP), C.c_ulonglong]),
    "milb_host_free": (C.c_int, [_VP]),
    "milb_dev_alloc": (C.c_int, [C.POINTER(_VP), C.c_ulonglong]),
    "milb_dev_free": (C.c_int, [_VP]),
    "milb_set_device": (C.c_int, [C.c_int]),
    "milb_memcpy": (C.c_int, [_VP, _VP, C.c_ulonglong]),
    "milb_ipc_export": (C.c_int, [_VP, C.c_char_p]),
    "milb_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(_VP)]),
    "milb_ipc_close": (C.c_int, [_VP]),
    "milb_reg_create": (C.c_int, [C.POINTER(_VP), _U]),
    "milb_reg_destroy": (None, [_VP]),
    "milb_reg_set_images": (C.c_int, [_VP, _VP, _VP, C.c_int, _VP]),
    "milb_reg_prepare": (C.c_int, [_VP, _F, _F, _VP]),
    "milb_reg_set_fetch": (C.c_int, [_VP, C.c_int]),
    "milb_reg_cost": (C.c_int, [_VP, _F, C.c_int, _F, _VP]),
    "milb_reg_cost_sums": (C.c_int, [_VP, _F, C.c_int, _D, _D, _VP]),
    "milb_reg_warp_source": (C.c_int, [_VP, _F, _VP, C.c_int, _VP]),
    "milb_affine_warp": (C.c_int, [_VP, _U, _VP, _U, _F, C.c_int, _VP]),
    "milb_reg3d_affine": (C.c_int, [_VP, _F, _VP, _VP, _U, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _F, _VP]),
    "milb_phasor": (C.c_int, [C.POINTER(_LL), _VP, _VP, _U, _VP]),
    "milb_imshift": (C.c_int, [_VP, _VP, _U, C.POINTER(_LL), _VP]),
    "milb_reg2d_create": (C.c_int, [C.POINTER(_VP), _VP, _U, _VP, _U, C.c_int, _F, _VP]),
    "milb_reg2d_destroy": (None, [_VP]),
    "milb_reg2d_cost": (C.c_int, [_VP, _F, C.c_int, _F, _VP]),
    "milb_reg2d_warp": (C.c_int, [_VP, _F, C.c_int, _VP, C.c_int, _VP]),
    "milb_reg2d_shiftalign": (C.c_int, [_VP, _F, _VP, _U, _VP, _U, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, _F, _VP]),
    "milb_reg2d_affine": (C.c_int, [_VP, _F, _VP, _U, _VP, _U, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _F, _VP]),
    "milb_p2matrix": (None, [_F, _F]),
    "milb_matrix2p": (None, [_F, _F]),
    "milb_matrixmultiply": (None, [_F, _F, _F]),
    "milb_dof9tomatrix": (None, [_F, _F, C.c_int]),
    "milb_powell": (C.c_int, [_F, _F, C.c_int, C.c_float, _I, _F, COSTFN, _VP, _I, C.c_int]),
}

_lib = None


def load() -> C.CDLL:
    """Returns the loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m microimagelib_b200.build` "
            "(or __graft_entry__.build()).  microimagelib_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for protos in (LIBAPI_PROTOS, CAPI_PROTOS):
        for name, (res, args) in protos.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            if args is not None:
                fn.argtypes = args
    _lib = lib
    return lib
