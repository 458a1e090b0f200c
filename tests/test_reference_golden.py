"""Committed vectors computed by the REFERENCE ITSELF on a B200 (tests/golden/reference_vectors.npz, written by
tests/golden/make_reference_golden.py from oracle/_ref/libapi_ref.so).  CPU: the oracle reproduces them (this
is what pins the oracle to the reference).  GPU: the product reproduces them through the C-ABI.
Tolerances are the north star's: deconvolution rel-L2 <= 1e-4, ZNCC <= 1e-5; projections / rotations bit-exact."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "reference_vectors.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/reference_vectors.npz not generated yet")


@pytest.fixture(scope="module")
def G():
    return dict(np.load(PATH))


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_oracle_reproduces_the_reference_decon(G):
    from oracle import decon_oracle as do
    a, b = G["img_a"], G["img_b"]
    assert rel_l2(do.decon_singleview(a, G["psf_a"], 5), G["decon_sv_5it"]) <= 1e-4
    assert rel_l2(do.decon_singleview(a, G["psf_even"], 4), G["decon_sv_even_psf_4it"]) <= 1e-4
    assert rel_l2(do.decon_singleview(a, G["psf_a"], 3, const_init=True), G["decon_sv_constinit_3it"]) <= 1e-4
    assert rel_l2(do.decon_singleview(a, G["psf_a"], 3, unmatch=True, psf_bp=G["psf_b"]), G["decon_sv_unmatched_3it"]) <= 1e-4
    assert rel_l2(do.decon_singleview(G["img_c"], G["psf_a"], 6), G["decon_sv_pow2_6it"]) <= 1e-4
    assert rel_l2(do.decon_dualview(a, b, G["psf_a"], G["psf_b"], 3), G["decon_dv_3it"]) <= 1e-4


def test_oracle_reproduces_the_reference_warp_and_zncc(G):
    from oracle import reg_oracle as ro
    a, m = G["img_a"], G["matrix"]
    w = ro.affine_warp(a, m)
    scale = float(np.abs(a).max())
    assert float(np.abs(w - G["warp"]).max()) <= 4e-6 * scale          # hardware filter vs its software restatement: a few ulp
    assert np.array_equal(w == 0, G["warp"] == 0)                       # same validity mask
    t_dm, sd = ro.demean(a)
    s_dm, _ = ro.demean(G["warp"])
    for k, want in zip(G["cost_matrices"], G["zncc"]):
        assert abs(-float(ro.zncc_cost(t_dm, sd, s_dm, k)) - float(want)) <= 1e-5


def test_oracle_reproduces_the_reference_geometry(G):
    from oracle import reg_oracle as ro
    a = G["img_a"]
    assert np.array_equal(ro.max_projection(a, 3) if False else a.max(axis=0), G["mip_z"])   # zero-initialised max of positive data
    assert np.array_equal(a.max(axis=2), G["mip_x"])
    assert np.array_equal(a.max(axis=1).T, G["mip_y"])
    sz, sy, sx = a.shape
    want = np.transpose(a, (2, 1, 0))[::-1]          # +90 deg about Y: out[x' = k, y' = j, z' = sx-1-i] = in[i, j, k] (cukernel.cuh:437-453)
    assert np.array_equal(want, G["rot_plus90"])


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference(G):
    from microimagelib_b200 import device, libapi
    a, b, m = G["img_a"], G["img_b"], G["matrix"]
    got, st, _ = libapi.decon_singleview(a, G["psf_a"], 5)
    assert st == 0 and rel_l2(got, G["decon_sv_5it"]) <= 1e-4
    got, st, _ = libapi.decon_singleview(a, G["psf_even"], 4)
    assert st == 0 and rel_l2(got, G["decon_sv_even_psf_4it"]) <= 1e-4
    got, st, _ = libapi.decon_singleview(a, G["psf_a"], 3, initialFlag=True)
    assert st == 0 and rel_l2(got, G["decon_sv_constinit_3it"]) <= 1e-4
    got, st, _ = libapi.decon_singleview(a, G["psf_a"], 3, flagUnmatch=True, psf_bp=G["psf_b"])
    assert st == 0 and rel_l2(got, G["decon_sv_unmatched_3it"]) <= 1e-4
    got, st, _ = libapi.decon_singleview(G["img_c"], G["psf_a"], 6)
    assert st == 0 and rel_l2(got, G["decon_sv_pow2_6it"]) <= 1e-4
    got, st, _ = libapi.decon_dualview(a, b, G["psf_a"], G["psf_b"], 3)
    assert st == 0 and rel_l2(got, G["decon_dv_3it"]) <= 1e-4
    w, st = libapi.atrans3dgpu(a, m)
    assert float(np.abs(w - G["warp"]).max()) <= 4e-6 * float(np.abs(a).max())
    for mode in ("hw", "sw"):
        r = device.Reg(a.shape, fetch=mode)
        r.set_images(a, G["warp"])
        r.prepare()
        z = -r.cost(G["cost_matrices"])
        r.close()
        assert float(np.abs(z - G["zncc"]).max()) <= 1e-5
    zp, xp, yp, st = libapi.mp2dgpu(a)
    assert np.array_equal(zp, G["mip_z"]) and np.array_equal(xp, G["mip_x"]) and np.array_equal(yp, G["mip_y"])
    rot, st = libapi.imoperation3D(a, 1)
    assert np.array_equal(rot, G["rot_plus90"])
    rs, st = libapi.imresize3d(a, (24, 24, 40))
    assert float(np.abs(rs - G["resize"]).max()) <= 4e-6 * float(np.abs(a).max())
    mp, st = libapi.mip3dgpu(a, 2, 4)
    assert mp.shape == G["mip3d_y"].shape and float(np.abs(mp - G["mip3d_y"]).max()) <= 4e-6 * float(np.abs(a).max())
    # whole registration with the default (hardware) fetch: the reference's trajectory
    _, tmx, st, rec = libapi.reg3d(a, G["warp"], regChoice=2, regMethod=7, FTOL=1e-4, itLimit=3000)
    assert st == 0 and abs(float(rec[1]) - float(G["reg3d_m7_records"][1])) <= 1e-5
    assert abs(float(rec[3]) - float(G["reg3d_m7_records"][3])) <= 1e-3
