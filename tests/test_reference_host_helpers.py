"""Pins the host-side helpers of the registration / deconvolution path to the reference's OWN compiled
functions (oracle/_ref/libapi_ref.so, built from /root/reference by oracle/build_ref_gpu.py; these are
plain C functions, no GPU needed): snapTransformSize, p2matrix, matrix2p, matrixmultiply, dof9tomatrix,
checkmatrix (src/api_subfunc.cu:57-87, 557-624, 715-824; src/api_reg.cpp:247-262).  Checked bit for bit:
  oracle  (oracle/reg_oracle.*, oracle/decon_oracle.py)   ==  reference
  product (lib/libapi.so: milb_* C-ABI, checkmatrix)      ==  reference
"""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_gpu

pytestmark = pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref/libapi_ref.so not built (needs /root/reference at build time)")

F = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def ref():
    lib = ref_gpu.load()
    lib._Z17snapTransformSizei.argtypes = [C.c_int]
    lib._Z17snapTransformSizei.restype = C.c_int
    for name in ("_Z8p2matrixPfS_", "_Z8matrix2pPfS_"):
        getattr(lib, name).argtypes = [F, F]
        getattr(lib, name).restype = None
    lib.matrixmultiply.argtypes = [F, F, F]
    lib.matrixmultiply.restype = None
    lib.dof9tomatrix.argtypes = [F, F, C.c_int]
    lib.dof9tomatrix.restype = None
    return lib


def _fp(a):
    return a.ctypes.data_as(F)


def test_snap_transform_size_table(ref):
    from microimagelib_b200 import _lib
    from oracle import decon_oracle as do
    prod = _lib.load()
    for n in list(range(1, 1400)) + [2047, 2048, 2049, 4000]:
        want = ref._Z17snapTransformSizei(n)
        assert do.snap_transform_size(n) == want, n
        assert prod.milb_snap_transform_size(n) == want, n


def test_parameter_matrix_maps(ref):
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    rng = np.random.default_rng(0)
    for _ in range(200):
        x = np.zeros(13, np.float32)
        x[1:] = rng.normal(0, 2, 12).astype(np.float32)
        m = np.zeros(12, np.float32)
        ref._Z8p2matrixPfS_(_fp(m), _fp(x))
        assert np.array_equal(ro.p2matrix(x), m) and np.array_equal(device.p2matrix(x), m)
        back = np.zeros(13, np.float32)
        ref._Z8matrix2pPfS_(_fp(m), _fp(back))
        assert np.array_equal(ro.matrix2p(m)[1:], back[1:]) and np.array_equal(device.matrix2p(m)[1:], back[1:])
        m2 = rng.normal(0, 1, 12).astype(np.float32)
        out = np.zeros(12, np.float32)
        ref.matrixmultiply(_fp(out), _fp(m), _fp(m2))
        assert np.array_equal(ro.matrixmultiply(m, m2), out)
        assert np.array_equal(device.matrixmultiply(m, m2), out)


@pytest.mark.parametrize("dof", [3, 6, 7, 9])
def test_dof9tomatrix(ref, dof):
    from microimagelib_b200 import device
    from oracle import reg_oracle as ro
    rng = np.random.default_rng(dof)
    for _ in range(300):
        q = np.zeros(10, np.float32)
        q[1:4] = rng.normal(0, 5, 3)
        q[4:7] = rng.normal(0, 20, 3)
        q[7:10] = 1 + rng.normal(0, 0.1, 3)
        q = q.astype(np.float32)
        m = np.zeros(12, np.float32)
        ref.dof9tomatrix(_fp(m), _fp(q), dof)
        assert np.array_equal(ro.dof9tomatrix(q, dof), m)
        assert np.array_equal(device.dof9tomatrix(q, dof), m)


def test_checkmatrix(ref):
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    rng = np.random.default_rng(7)
    R = ref_gpu.api()
    for _ in range(500):
        m = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
        m[[0, 5, 10]] = rng.uniform(0.3, 1.6, 3)
        m[[3, 7, 11]] = rng.uniform(-120, 120, 3)
        want = R.checkmatrix(m, 128, 96, 64)
        assert ro.checkmatrix(m, 128, 96, 64) == want
        assert libapi.checkmatrix(m, 128, 96, 64) == want
