"""CPU emulation of the device FFT stages (csrc/fft_core.h + fft_plan.h compiled with g++):
the mixed-radix butterflies, twiddles, position tables and the two-real-pencils-in-one-complex
split/merge are checked against numpy's float64 FFT for every size class snapTransformSize yields."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emul", "libfft_emul.so")
F = C.POINTER(C.c_float)
I = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "emul", "fft_emul.cpp")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", SO, src], check=True)
    return C.CDLL(SO)


SIZES = [16, 32, 64, 128, 192, 256, 320, 384, 448, 512, 576, 640, 704, 768, 832, 896, 960, 1024, 1088, 1216, 2048]


@pytest.mark.parametrize("n", SIZES)
def test_forward_inverse_and_positions(lib, n):
    rng = np.random.default_rng(n)
    radix = (C.c_int * 8)()
    pos = np.zeros(n, np.int32)
    ns = lib.emul_plan(n, radix, pos.ctypes.data_as(I))
    assert ns > 0 and int(np.prod([radix[i] for i in range(ns)])) == n
    assert sorted(pos.tolist()) == list(range(n))
    x = (rng.standard_normal((n, 4)) + 1j * rng.standard_normal((n, 4))).astype(np.complex64)
    d = x.copy()
    assert lib.emul_fft(n, 4, d.ctypes.data_as(F), 0) == 0
    ref = np.fft.fft(x.astype(np.complex128), axis=0)
    assert np.linalg.norm(d[pos] - ref) / np.linalg.norm(ref) < 5e-7
    assert lib.emul_fft(n, 4, d.ctypes.data_as(F), 1) == 0
    assert np.linalg.norm(d / n - x) / np.linalg.norm(x) < 5e-7


@pytest.mark.parametrize("n", SIZES)
def test_two_real_pencils_in_one_complex(lib, n):
    rng = np.random.default_rng(n + 1)
    a = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    A = np.zeros(n // 2 + 1, np.complex64)
    B = np.zeros(n // 2 + 1, np.complex64)
    lib.emul_r2c_pair(n, a.ctypes.data_as(F), b.ctypes.data_as(F), A.ctypes.data_as(F), B.ctypes.data_as(F))
    for got, src in ((A, a), (B, b)):
        ref = np.fft.rfft(src.astype(np.float64))
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 5e-7
    a2 = np.zeros(n, np.float32)
    b2 = np.zeros(n, np.float32)
    lib.emul_c2r_pair(n, A.ctypes.data_as(F), B.ctypes.data_as(F), a2.ctypes.data_as(F), b2.ctypes.data_as(F))
    assert np.abs(a2 / n - a).max() < 5e-6 and np.abs(b2 / n - b).max() < 5e-6


def test_unsupported_prime_factor_is_rejected(lib):
    radix = (C.c_int * 8)()
    assert lib.emul_plan(64 * 67, radix, None) == -1


@pytest.mark.parametrize("r", [2, 4, 8, 16, 32])
def test_register_butterflies_are_dfts(lib, r):
    """bfly2/4/8/16/32 (fft_core.h), the in-register transforms of the fast kernels' stages: natural order in
    and out, forward = DFT, inverse = un-normalised inverse DFT."""
    rng = np.random.default_rng(100 + r)
    x = (rng.standard_normal(r) + 1j * rng.standard_normal(r)).astype(np.complex64)
    for inverse in (0, 1):
        d = x.copy()
        assert lib.emul_bfly(r, d.ctypes.data_as(F), inverse) == 0
        ref = np.fft.ifft(x.astype(np.complex128)) * r if inverse else np.fft.fft(x.astype(np.complex128))
        assert np.linalg.norm(d - ref) / np.linalg.norm(ref) < 3e-7


@pytest.mark.parametrize("r", [3, 5, 7])
def test_odd_radix_register_butterflies(lib, r):
    """bfly3 / bfly5 / bfly7 (closed forms with constants, the last stage of the 64*k fast plans) == numpy's DFT"""
    rng = np.random.default_rng(r)
    for inverse in (0, 1):
        x = (rng.standard_normal(r) + 1j * rng.standard_normal(r)).astype(np.complex64)
        d = x.copy()
        assert lib.emul_bfly_odd(r, d.ctypes.data_as(F), inverse) == 0
        ref = np.fft.ifft(x.astype(np.complex128)) * r if inverse else np.fft.fft(x.astype(np.complex128))
        assert np.abs(d - ref).max() <= 1e-6 * np.abs(ref).max()


ZROW = {64: (8, 8), 128: (8, 16), 256: (16, 16), 512: (16, 32), 1024: (32, 32)}


@pytest.mark.parametrize("n", sorted(ZROW))
def test_row_convolution_lanes(lib, n):
    """k_zrow's per-lane stage code (csrc/zrow_core.h) replayed lane by lane: the forward-only mode yields numpy's FFT in the
    kernel's per-row OTF order, the convolution mode with that OTF equals the circular convolution, and no 16-lane group of a
    shared-memory access hits a bank twice."""
    r0, r1 = ZROW[n]
    rng = np.random.default_rng(n + 7)
    freq = np.zeros(n, np.int32)
    assert lib.emul_zrow_maps(n, r0, r1, freq.ctypes.data_as(I)) == 0
    assert sorted(freq.tolist()) == list(range(n))
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    k = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    conflicts = C.c_int(-1)
    K = k.copy()
    assert lib.emul_zrow(n, K.ctypes.data_as(F), None, 0, C.c_float(1.0 / n), C.byref(conflicts)) == 0
    assert conflicts.value == 0
    ref_K = np.fft.fft(k.astype(np.complex128)) / n
    assert np.linalg.norm(K - ref_K[freq]) / np.linalg.norm(ref_K) < 5e-7
    d = x.copy()
    assert lib.emul_zrow(n, d.ctypes.data_as(F), K.ctypes.data_as(F), 1, C.c_float(1.0), C.byref(conflicts)) == 0
    ref = np.fft.ifft(np.fft.fft(x.astype(np.complex128)) * np.fft.fft(k.astype(np.complex128)))
    assert np.linalg.norm(d - ref) / np.linalg.norm(ref) < 1e-6


@pytest.mark.parametrize("r0,r1", [(16, 32), (16, 16), (8, 16)])
def test_folded_x_pass_index_rules(lib, r0, r1):
    """xfold_core.h: every slot of the radix-r1 stages builds the frequencies k1 + r0 * k2 of its row from the half spectrum
    (directly for k2 < r1 / 2, as mirrors N - k otherwise), mirror partners sit in adjacent slots (the two halves of one warp
    at 16 lanes), and the split covers every row 0 .. N / 2 of the half spectrum exactly once."""
    n = r0 * r1
    row_of = np.zeros(r0 * r1, np.int32)
    mirror = np.zeros(r0 * r1, np.int32)
    rows = np.zeros(r0, np.int32)
    assert lib.emul_xfold_inputs(r0, r1, row_of.ctypes.data_as(I), mirror.ctypes.data_as(I), rows.ctypes.data_as(I)) == 0
    assert sorted(rows.tolist()) == list(range(r0))
    rng = np.random.default_rng(r0 * r1)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    A, B, Cs = np.fft.fft(a), np.fft.fft(b), np.fft.fft(a + 1j * b)
    for s in range(r0):
        k1 = int(rows[s])
        assert int(rows[s ^ 1]) == ((r0 - k1) % r0 if k1 not in (0, r0 // 2) else (r0 // 2 if k1 == 0 else 0))
        for k2 in range(r1):
            k = int(row_of[s * r1 + k2])
            assert 0 <= k <= n // 2
            ck, cn = A[k] + 1j * B[k], np.conj(A[k]) + 1j * np.conj(B[k])   # merge_pair of half-spectrum row k
            got = cn if mirror[s * r1 + k2] else ck
            assert abs(got - Cs[k1 + r0 * k2]) < 1e-9 * n
    # split: slot pairs (k1, r0 - k1) output rows k1 + r0 * k2, k2 < r1 / 2; row 0 also k2 = r1 / 2
    out = [k1 + r0 * k2 for k1 in range(r0) for k2 in range(r1 // 2)] + [n // 2]
    assert sorted(out) == list(range(n // 2 + 1))
