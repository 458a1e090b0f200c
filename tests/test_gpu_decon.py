"""GPU parity of the Richardson-Lucy path against the CPU oracle (oracle/decon_oracle.py),
called through the C-ABI (libapi.so).  Tolerance from BASELINE.json north_star:
relative L2 error <= 1e-4 after the stated iteration count."""
import numpy as np
import pytest

from microimagelib_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-4  # north_star: "deconvolved volumes within a relative L2 error of 1e-4"


def rel_l2(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _case(shape, psf_shape, sigma, seed=synth.SEED_A):
    psf = synth.gaussian_psf(psf_shape, sigma)
    img = synth.bead_image(shape, psf, seed=seed)
    return img, psf


def test_config1_singleview_128cube_10it():
    """BASELINE config 1: 128x128x128 beads + Gaussian PSF, 10 iterations."""
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    img, psf = _case((128, 128, 128), (65, 65, 65), (4, 2, 2))
    got, st, rec = libapi.decon_singleview(img, psf, 10)
    assert st == 0 and rec[0] == 1
    ref = do.decon_singleview(img, psf, 10)
    assert rel_l2(got, ref) <= TOL
    # flux is conserved by RL with a unit-DC OTF (SURVEY 8(c) KAT 2)
    assert abs(got.sum(dtype=np.float64) / np.maximum(img, 0.01).sum(dtype=np.float64) - 1) < 1e-3


@pytest.mark.parametrize("shape,psf_shape", [
    ((64, 64, 64), (33, 33, 33)),          # pow2 box, odd PSF
    ((40, 100, 150), (31, 33, 35)),        # box 64 x 128 x 192: padding + radix-3 axis
    ((96, 140, 90), (32, 32, 32)),         # even PSF (flip quirk), box 128 x 192 x 128
    ((20, 300, 36), (21, 21, 21)),         # box 32 x 320 x 64: radix-5 axis
    ((16, 16, 16), (33, 17, 9)),           # PSF larger than the box in one axis
])
def test_singleview_shapes(shape, psf_shape):
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    img, psf = _case(shape, psf_shape, (2.5, 2.0, 1.5))
    got, st, _ = libapi.decon_singleview(img, psf, 5)
    assert st == 0
    ref = do.decon_singleview(img, psf, 5)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) <= TOL


def test_singleview_const_init_and_unmatched():
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    img, psf = _case((48, 64, 80), (25, 25, 25), (3, 2, 2))
    bp = synth.gaussian_psf((25, 25, 25), (1.5, 1.0, 1.0))
    got, st, _ = libapi.decon_singleview(img, psf, 4, initialFlag=True)
    ref = do.decon_singleview(img, psf, 4, const_init=True)
    assert rel_l2(got, ref) <= TOL
    got, st, _ = libapi.decon_singleview(img, psf, 4, flagUnmatch=True, psf_bp=bp)
    ref = do.decon_singleview(img, psf, 4, unmatch=True, psf_bp=bp)
    assert rel_l2(got, ref) <= TOL


def test_dualview():
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    shape = (64, 96, 128)
    psf_a = synth.gaussian_psf((33, 33, 33), (4, 2, 2))
    psf_b = synth.gaussian_psf((33, 33, 33), (2, 2, 4))
    a = synth.bead_image(shape, psf_a, seed=synth.SEED_A, noise_seed=synth.SEED_NOISE)
    b = synth.bead_image(shape, psf_b, seed=synth.SEED_A, noise_seed=synth.SEED_NOISE + 1)
    got, st, rec = libapi.decon_dualview(a, b, psf_a, psf_b, 6)
    assert st == 0
    ref = do.decon_dualview(a, b, psf_a, psf_b, 6)
    assert rel_l2(got, ref) <= TOL
    got, st, rec = libapi.decon_dualview(a, b, psf_a, psf_b, 3, initialFlag=True)
    ref = do.decon_dualview(a, b, psf_a, psf_b, 3, const_init=True)
    assert rel_l2(got, ref) <= TOL


def test_delta_image_closed_form():
    """KAT: a constant image is a fixed point; a delta image after one iteration has the closed
    form E1 = max(A * (h_bp conv (A / (h conv A))), .01) -- checked against float64 numpy."""
    from microimagelib_b200 import libapi
    import scipy.fft as sfft
    shape = (32, 32, 32)
    psf = synth.gaussian_psf((15, 15, 15), (2, 2, 2))
    const = np.full(shape, 7.0, np.float32)
    got, st, _ = libapi.decon_singleview(const, psf, 3)
    assert np.allclose(got, 7.0, rtol=1e-5)
    img = np.full(shape, 1.0, np.float32)
    img[16, 16, 16] = 1000.0
    got, st, _ = libapi.decon_singleview(img, psf, 1)
    box = np.zeros(shape)
    idx = [(np.arange(15) - 7) % 32] * 3
    box[np.ix_(*idx)] = psf.astype(np.float64) / psf.astype(np.float64).sum()
    H = sfft.fftn(box)
    A = img.astype(np.float64)
    conv = sfft.ifftn(sfft.fftn(A) * H).real
    ratio = A / conv
    back = sfft.ifftn(sfft.fftn(ratio) * np.conj(H)).real  # odd symmetric PSF: flipped == conj
    want = np.maximum(A * back, 0.01)
    assert rel_l2(got, want) <= TOL


def test_bad_mode_returns_like_reference():
    from microimagelib_b200 import libapi
    img, psf = _case((16, 16, 16), (9, 9, 9), (1, 1, 1))
    _, st, _ = libapi.decon_singleview(img, psf, 1, gpuMemMode=7)
    assert st == 1      # src/api_decon.cpp:318
    _, st, _ = libapi.decon_dualview(img, img, psf, psf, 1, gpuMemMode=7)
    assert st == -1     # src/api_decon.cpp:687
    assert libapi.fusion_dualview_status() == 1   # src/api_decon.cpp:1133-1136


def test_device_resident_handle_matches_host_api():
    import torch
    from microimagelib_b200 import device, libapi
    img, psf = _case((64, 64, 64), (17, 17, 17), (2, 2, 2))
    ref, _, _ = libapi.decon_singleview(img, psf, 3)
    d = device.Decon(img.shape, 1)
    d.set_psf(0, psf)
    d.set_image(0, torch.from_numpy(img).cuda())
    d.run(3)
    out = torch.empty(img.shape, dtype=torch.float32, device="cuda")
    d.result(out)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    d.close()


FAST_SHAPES = [(64, 128, 256), (256, 64, 128), (128, 512, 64), (64, 64, 1024), (512, 64, 128), (1024, 64, 64), (128, 256, 512),
               (64, 1024, 64), (64, 1024, 1024),   # Y = 1024: k_ypassW (16-lane single-buffer tiles) next to the row convolution
               # the 64*k lengths of the compile-time plans, each in every axis role (X pencils / Y planes / Z planes)
               (192, 320, 64), (320, 64, 192), (64, 192, 320), (384, 448, 64), (448, 64, 384), (64, 384, 448),
               (576, 64, 640), (64, 640, 576), (640, 576, 64), (768, 64, 192), (64, 768, 128), (128, 64, 768)]


@pytest.mark.parametrize("shape", FAST_SHAPES)
def test_fast_pow2_path_matches_generic_path(shape, monkeypatch):
    """The compile-time fast kernels (fft_fast.cuh: power-of-two lengths and the 64*k lengths 192 ... 768) and the
    generic mixed-radix kernels (fft_kernels.cuh) are independent implementations of the same loop."""
    from microimagelib_b200 import device
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
    img = synth.bead_image(shape, psf, density=1 / 4096.0)
    outs = {}
    for name, env in (("fast", "0"), ("generic", "1")):
        monkeypatch.setenv("MILB_FORCE_GENERIC", env)
        d = device.Decon(shape, 1)
        d.set_psf(0, psf)
        d.set_image(0, img)
        d.run(6)
        outs[name] = d.result().copy()
        if name == "fast":
            for chunk in (1, 5, 0):          # L2 chunking must not change the result
                d.set_chunk_planes(chunk)
                d.run(6)
                assert np.array_equal(d.result(), outs["fast"])
        d.close()
    assert rel_l2(outs["fast"], outs["generic"]) <= 2e-6


def test_non_pow2_fast_box_against_the_oracle():
    """A box as snapTransformSize really produces it (300 x 400 x 90 image -> 320 x 448 x 96... here 320 x 448 x 192): the
    compile-time 64*k plans against the CPU oracle, padding and cropping included."""
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    psf = synth.gaussian_psf((17, 17, 17), (2.0, 2.0, 2.5))
    img = synth.bead_image((180, 400, 300), psf, density=1 / 4096.0)
    assert do.fft_shape_for(img.shape) == (192, 448, 320)
    got, st, _ = libapi.decon_singleview(img, psf, 5)
    ref = do.decon_singleview(img, psf, 5)
    assert st == 0 and rel_l2(got, ref) <= 1e-4


def test_edge_inputs_zero_iterations_clamp_and_padded_dualview():
    """zero iterations return the (clamped) initial estimate; zeros and negative voxels are raised to 0.01
    (src/api_subfunc.cu:3380); dual view on a padded, non-power-of-two box with unmatched back projectors."""
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    img, psf = _case((24, 40, 56), (13, 13, 13), (2, 2, 2))
    img = img - np.float32(np.percentile(img, 60))                 # 60 % of the voxels are now <= 0
    got, st, _ = libapi.decon_singleview(img, psf, 0)
    assert st == 0 and np.array_equal(got, np.maximum(img, np.float32(0.01)))
    assert np.array_equal(got, do.decon_singleview(img, psf, 0))
    got, st, _ = libapi.decon_singleview(img, psf, 4)
    assert rel_l2(got, do.decon_singleview(img, psf, 4)) <= TOL and got.min() >= np.float32(0.01)
    shape = (40, 100, 72)                                           # box 64 x 128 x 128 after padding
    pa, pb = synth.gaussian_psf((21, 21, 21), (3, 2, 2)), synth.gaussian_psf((21, 21, 21), (2, 2, 3))
    ba, bb = synth.gaussian_psf((21, 21, 21), (1.5, 1, 1)), synth.gaussian_psf((21, 21, 21), (1, 1, 1.5))
    a = synth.bead_image(shape, pa, density=1 / 2048.0, seed=7)
    b = synth.bead_image(shape, pb, density=1 / 2048.0, seed=7, noise_seed=8)
    got, st, _ = libapi.decon_dualview(a, b, pa, pb, 4, flagUnmatch=True, psf_bp1=ba, psf_bp2=bb)
    ref = do.decon_dualview(a, b, pa, pb, 4, unmatch=True, psf_bp1=ba, psf_bp2=bb)
    assert st == 0 and rel_l2(got, ref) <= TOL


@pytest.mark.parametrize("shape,dual", [((64, 128, 128), False), ((128, 64, 256), True), ((64, 192, 320), False)])
def test_pipelined_host_copies_give_the_same_volume(shape, dual, monkeypatch):
    """decon_singleview / decon_dualview on host images of exactly the FFT box size overlap the H2D / D2H copies with the
    first and the last X pass (milb_decon_run_host, row-range chunks): bit-identical to the plain upload -> loop -> download."""
    from microimagelib_b200 import libapi
    psf_a = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
    psf_b = synth.gaussian_psf((17, 17, 17), (1.5, 2.0, 2.5))
    a = synth.bead_image(shape, psf_a, density=1 / 4096.0)
    b = synth.bead_image(shape, psf_b, density=1 / 4096.0, noise_seed=5)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MILB_HOST_PIPELINE", mode)
        for rep in range(2):                                   # the second call reuses the cached handle and its copy stream
            if dual:
                out, st, _ = libapi.decon_dualview(a, b, psf_a, psf_b, 4)
            else:
                out, st, _ = libapi.decon_singleview(a, psf_a, 4)
            assert st == 0
        outs[mode] = out.copy()
    assert np.array_equal(outs["0"], outs["1"])


@pytest.mark.parametrize("shape,dual", [((64, 64, 64), False), ((64, 128, 128), True), ((128, 64, 256), False), ((64, 192, 512), False),
                                        ((64, 320, 128), True), ((64, 64, 1024), False), ((64, 1024, 128), False)])
def test_row_convolution_equals_the_transposing_plane_kernels(shape, dual):
    """k_zrow (Z convolution along the contiguous axis, warp-private pencils, OTFs in its per-row order) runs the same
    butterflies in the same order as k_ypassT + k_zconvT: bit-identical volumes for the lengths that share FastPlan's
    two-stage plan; Z = 1024 uses the row kernel's own 32 x 32 plan (the transposing kernels run 8 x 8 x 4 x 4) and agrees
    to rounding."""
    from microimagelib_b200 import device
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
    img = synth.bead_image(shape, psf, density=1 / 4096.0)
    outs = []
    for row in (True, False):
        d = device.Decon(shape, 2 if dual else 1, row_conv=row)
        for v in range(2 if dual else 1):
            d.set_psf(v, psf)
            d.set_image(v, img)
        d.run(5)
        outs.append(d.result().copy())
        d.close()
    if 1024 in shape[1:]:   # Z = 1024: the row kernel's 32 x 32 plan; Y = 1024: k_ypassW's, against 8 x 8 x 4 x 4
        assert rel_l2(outs[0], outs[1]) <= 1e-6
    else:
        assert np.array_equal(outs[0], outs[1])


def test_plane_pipeline_equals_the_three_launches(monkeypatch):
    """MILB_PLANE_PIPE=1 (Y forward | row convolution | Y inverse side by side on disjoint SMs, planes handed over through
    per-plane counters): an experiment that is off by default, but it must stay bit-identical to the three launches and must
    not hang (its polls trap)."""
    from microimagelib_b200 import device
    shape = (64, 256, 256)
    psf = synth.gaussian_psf((17, 17, 17), (2.5, 2.0, 1.5))
    img = synth.bead_image(shape, psf, density=1 / 4096.0)
    outs = []
    for pipe in ("1", "0"):
        monkeypatch.setenv("MILB_PLANE_PIPE", pipe)
        d = device.Decon(shape, 1)
        assert d.row_convolution()
        d.set_psf(0, psf)
        d.set_image(0, img)
        d.run(4)
        outs.append(d.result().copy())
        if pipe == "1":
            ms = d.time_pipe(2)
            assert ms.shape == (4,) and (ms > 0).all()
        d.close()
    assert np.array_equal(outs[0], outs[1])
