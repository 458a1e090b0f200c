"""ctypes binding of oracle/_ref/libapi_ref.so -- the REFERENCE's own libapi, built from its
unmodified algorithms by oracle/build_ref_gpu.py (mechanical texture-object patch only) and run on the
GPU.  TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_reference_golden.py and bench.py's
reference-structure yardstick, never by the product.

`api` exposes the same Python mirror of include/libapi.h as microimagelib_b200.libapi (same functions,
same argument conventions -- the mirror's source is re-executed against this library), so a parity test
reads   ref_gpu.api().decon_singleview(img, psf, 10)   vs   libapi.decon_singleview(img, psf, 10).
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import types

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libapi_ref.so")

_lib = None
_api = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load():
    """The reference library with the prototypes of include/libapi.h declared."""
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run python oracle/build_ref_gpu.py where /root/reference exists")
        from microimagelib_b200._lib import LIBAPI_PROTOS
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in LIBAPI_PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            if args is not None:
                fn.argtypes = args
        # internals of the reference used by the pinned tests (C++ symbols have C linkage only where the
        # reference declares them extern "C"; the rest are reached through the public API)
        _lib = lib
    return _lib


def api():
    """microimagelib_b200/libapi.py re-executed with the reference library behind it."""
    global _api
    if _api is None:
        spec = importlib.util.find_spec("microimagelib_b200.libapi")
        src = open(spec.origin).read().replace("from . import _lib", "")
        mod = types.ModuleType("oracle._ref_api")
        shim = types.SimpleNamespace(load=load)
        mod.__dict__["_lib"] = shim
        exec(compile(src, spec.origin + " [reference library]", "exec"), mod.__dict__)
        _api = mod
    return _api
