"""Pins the CPU oracle of the RL path (oracle/decon_oracle.py) with known-answer tests.  The
reference ships no golden vectors, so these are the KATs of SURVEY 8(c):
snapTransformSize table, OTF unit DC, odd PSF => OTF_bp == conj(OTF), even PSF => one-voxel shift,
closed-form first iteration, flux conservation, float32-vs-float64 agreement."""
import numpy as np
import scipy.fft as sfft

from microimagelib_b200 import synth
from oracle import decon_oracle as do


def test_snap_transform_size_table():
    # src/api_subfunc.cu:57-87 -- values worked by hand from the rule
    table = {1: 16, 16: 16, 17: 32, 33: 64, 64: 64, 65: 128, 100: 128, 128: 128, 129: 192, 140: 192, 192: 192,
             193: 256, 256: 256, 257: 320, 300: 320, 321: 384, 500: 512, 512: 512, 513: 576, 1000: 1024, 1024: 1024,
             1025: 1088, 2000: 2048}
    for n, want in table.items():
        assert do.snap_transform_size(n) == want, n


def test_pad_crop_roundtrip_and_edge_replication():
    rng = np.random.default_rng(0)
    img = rng.random((5, 6, 7)).astype(np.float32)
    box = do.pad_stack(img, (16, 16, 16))
    assert np.array_equal(do.crop_stack(box, img.shape), img)
    o = [(16 - s) // 2 for s in img.shape]
    assert box[0, 0, 0] == img[0, 0, 0] and box[-1, -1, -1] == img[-1, -1, -1]
    assert np.array_equal(box[o[0] + 2, :o[1], o[2] + 3], np.full(o[1], img[2, 0, 3]))


def test_pad_psf_centre_goes_to_origin():
    psf = np.zeros((5, 4, 3), np.float32)
    psf[2, 2, 1] = 1       # floor(P/2) per axis
    psf[0, 0, 0] = 2
    out = do.pad_psf(psf, (8, 8, 8))
    assert out[0, 0, 0] == 1 and out[-2, -2, -1] == 2 and out.sum() == 3


def test_align_size_truncating_division():
    v = np.arange(5 * 4 * 3, dtype=np.float32).reshape(5, 4, 3)
    out = do.align_size(v, (4, 6, 3))      # (4-5)/2 -> 0 in C, (6-4)/2 -> 1
    assert np.array_equal(out[:, 1:5, :], v[:4])
    assert out[:, 0].sum() == 0 and out[:, 5].sum() == 0


def test_otf_unit_dc_and_matched_backprojector():
    psf = synth.gaussian_psf((17, 15, 13), (2, 1.5, 1))
    otf, otf_bp = do.gen_otf_pair(psf, (32, 32, 16))
    assert abs(otf[0, 0, 0] - 1) < 1e-6
    assert np.abs(otf_bp - np.conj(otf)).max() < 5e-7          # odd, symmetric PSF (KAT 6)
    even = synth.gaussian_psf((16, 16, 16), (2, 2, 2))
    otf, otf_bp = do.gen_otf_pair(even, (32, 32, 32))
    # even size: flipping moves the centre by one voxel => OTF_bp = conj(OTF) * phase ramp
    k = [np.fft.fftfreq(32)[:, None, None], np.fft.fftfreq(32)[None, :, None], np.fft.rfftfreq(32)[None, None, :]]
    # PSF grid is symmetric about index 8 (= floor(16/2)); flipped it is symmetric about 7
    ramp = np.exp(2j * np.pi * (k[0] + k[1] + k[2]))
    assert np.abs(otf_bp - np.conj(otf) * ramp).max() < 5e-6 or np.abs(otf_bp - np.conj(otf) * np.conj(ramp)).max() < 5e-6
    assert np.abs(otf_bp - np.conj(otf)).max() > 1e-3


def test_psf_larger_than_box_is_cropped_then_shifted():
    psf = synth.gaussian_psf((33, 9, 9), (3, 1, 1))
    otf = do.gen_otf(psf, (16, 16, 16))
    boxed = do.align_size(psf * np.float32(1.0 / psf.sum(dtype=np.float64)), (16, 16, 16))
    want = sfft.rfftn(np.roll(boxed.astype(np.float64), (-8, -8, -8), axis=(0, 1, 2)))
    assert np.abs(otf - want).max() < 1e-6


def test_first_iteration_closed_form_and_fixed_point():
    shape = (16, 32, 32)
    psf = synth.gaussian_psf((9, 9, 9), (1.5, 1.5, 1.5))
    const = np.full(shape, 5.0, np.float32)
    assert np.allclose(do.decon_singleview(const, psf, 4), 5.0, rtol=2e-6)
    img = np.full(shape, 1.0, np.float32)
    img[8, 16, 16] = 500.0
    got = do.decon_singleview(img, psf, 1)
    box = np.zeros(shape)
    box[np.ix_((np.arange(9) - 4) % 16, (np.arange(9) - 4) % 32, (np.arange(9) - 4) % 32)] = psf.astype(np.float64)
    box /= box.sum()
    H = sfft.fftn(box)
    A = img.astype(np.float64)
    ratio = A / sfft.ifftn(sfft.fftn(A) * H).real
    want = np.maximum(A * sfft.ifftn(sfft.fftn(ratio) * np.conj(H)).real, 0.01)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5


def test_flux_conservation_and_float64_agreement():
    psf = synth.gaussian_psf((17, 17, 17), (3, 2, 2))
    img = synth.bead_image((32, 64, 64), psf, density=1 / 2048.0)   # == its FFT box: no pad / crop
    out = do.decon_singleview(img, psf, 10)
    assert abs(out.sum(dtype=np.float64) / img.sum(dtype=np.float64) - 1) < 1e-5      # KAT 2
    # float64 re-statement of the same loop
    fshape = do.fft_shape_for(img.shape)
    A = do.pad_stack(img, fshape).astype(np.float64)
    otf, otf_bp = [o.astype(np.complex128) for o in do.gen_otf_pair(psf, fshape)]
    E = A.copy()
    for _ in range(10):
        T = sfft.irfftn(sfft.rfftn(E) * otf, s=fshape)
        T = sfft.irfftn(sfft.rfftn(A / T) * otf_bp, s=fshape)
        E = np.maximum(E * T, 0.01)
    E = do.crop_stack(E, img.shape)
    assert np.linalg.norm(out - E) / np.linalg.norm(E) < 1e-5


def test_const_init_uses_the_sum_not_the_mean():
    # src/api_subfunc.cu:3382 (sic)
    psf = synth.gaussian_psf((5, 5, 5), (1, 1, 1))
    img = np.full((16, 16, 16), 2.0, np.float32)
    otf, otf_bp = do.gen_otf_pair(psf, (16, 16, 16))
    out = do.rl_single(img, otf, otf_bp, 0, const_init=True)
    assert np.all(out == np.float32(2.0 * 16 ** 3))


def test_dualview_init_and_alternation():
    psf_a = synth.gaussian_psf((9, 9, 9), (2, 1, 1))
    psf_b = synth.gaussian_psf((9, 9, 9), (1, 1, 2))
    a = synth.bead_image((16, 32, 32), psf_a, density=1 / 1024.0)
    b = synth.bead_image((16, 32, 32), psf_b, density=1 / 1024.0, noise_seed=7)
    o1, b1 = do.gen_otf_pair(psf_a, a.shape)
    o2, b2 = do.gen_otf_pair(psf_b, a.shape)
    assert np.array_equal(do.rl_dual(a, b, o1, b1, o2, b2, 0), (np.maximum(a, .01) + np.maximum(b, .01)) * np.float32(0.5))
    one = do.rl_dual(a, b, o1, b1, o2, b2, 1)
    E = (np.maximum(a, .01) + np.maximum(b, .01)) * np.float32(0.5)
    E = do._rl_half_step(E, np.maximum(a, np.float32(.01)), o1, b1)
    E = do._rl_half_step(E, np.maximum(b, np.float32(.01)), o2, b2)
    assert np.array_equal(one, E)
