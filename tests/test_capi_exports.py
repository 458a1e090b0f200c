"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//.*", "", txt)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", txt):
        n = m.group(1)
        if n in ("defined", "sizeof", "float", "int", "void", "char", "double", "long", "short", "unsigned") or n.startswith("MILB_"):
            continue
        names.add(n)
    # function-pointer typedef names are types, not symbols
    names -= set(re.findall(r"\(\s*\*\s*([A-Za-z_][A-Za-z0-9_]*)\s*\)", txt))
    return names


def test_library_loads_and_exports_declared_symbols():
    from microimagelib_b200 import _lib
    lib = _lib.load()
    api = _declared("libapi.h")
    capi = _declared("milb_capi.h")
    assert len(api) == 23, sorted(api)       # the reference's 23 entry points (include/libapi.h:12-68)
    for name in sorted(api | capi):
        assert hasattr(lib, name), name
    assert set(_lib.LIBAPI_PROTOS) == api
    assert set(_lib.CAPI_PROTOS) == capi


def test_host_only_helpers_work_without_gpu():
    from microimagelib_b200 import _lib, device
    lib = _lib.load()
    assert lib.milb_version().startswith(b"microimagelib_b200")
    assert [lib.milb_snap_transform_size(n) for n in (100, 300, 1000, 129)] == [128, 320, 1024, 192]
    import numpy as np
    from oracle import reg_oracle as ro
    rng = np.random.default_rng(0)
    for dof in (3, 6, 7, 9):
        q = np.concatenate([[0], rng.uniform(-20, 20, 6), rng.uniform(0.8, 1.2, 3)]).astype(np.float32)
        assert np.array_equal(device.dof9tomatrix(q, dof), ro.dof9tomatrix(q, dof))
    x = rng.standard_normal(13).astype(np.float32)
    assert np.array_equal(device.p2matrix(x), ro.p2matrix(x))
    m = rng.standard_normal(12).astype(np.float32)
    m2 = rng.standard_normal(12).astype(np.float32)
    assert np.array_equal(device.matrix2p(m), ro.matrix2p(m))
    assert np.array_equal(device.matrixmultiply(m, m2), ro.matrixmultiply(m, m2))
    from microimagelib_b200 import libapi
    assert libapi.checkmatrix([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], 10, 10, 10)
    assert not libapi.checkmatrix([1, 0, 0, 9, 0, 1, 0, 0, 0, 0, 1, 0], 10, 10, 10)


def test_volumes_beyond_32_bit_indexing_are_refused():
    """The warp / cost kernels index voxels with 32-bit integers: a 2048 x 2048 x 512 volume (2^31 voxels) fits a 180 GB GPU
    but must be refused, not silently mis-indexed (checked before any CUDA call, so this runs without a GPU)."""
    import ctypes as C
    from microimagelib_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    size = (C.c_uint * 3)(2048, 2048, 512)
    assert lib.milb_reg_create(C.byref(h), size) == 3          # MILB_ERR_SIZE
    assert not h.value
