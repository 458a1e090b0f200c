"""The committed hot-path vectors (tests/golden/path_vectors.npz, written by make_path_golden.py): the CPU
oracle must still reproduce them (CPU test), and the CUDA path must match them through the C-ABI (GPU test)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
G = np.load(os.path.join(HERE, "golden", "path_vectors.npz"))


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))


def test_oracle_reproduces_the_committed_vectors():
    import make_path_golden as mk
    now = mk.compute()
    for k in ("img_a", "img_b", "psf_a", "psf_b", "matrix"):
        assert np.array_equal(now[k], G[k]), k                       # the inputs are deterministic
    assert rel_l2(now["decon_sv_5it"], G["decon_sv_5it"]) <= 1e-6     # pocketfft versions may differ in the last bits
    assert rel_l2(now["decon_dv_3it"], G["decon_dv_3it"]) <= 1e-6
    assert np.array_equal(now["warp"], G["warp"])                     # integer-weight trilinear fetch: bit-exact
    assert np.array_equal(now["costs"], G["costs"])
    assert np.array_equal(now["phasor_shift"], G["phasor_shift"]) and list(G["phasor_shift"]) == [3, -2, 1]
    assert np.array_equal(now["mip_prealign"], G["mip_prealign"])


@pytest.mark.gpu
def test_cuda_path_matches_the_committed_vectors():
    from microimagelib_b200 import device, libapi
    from oracle import reg_oracle as ro
    got, st, _ = libapi.decon_singleview(G["img_a"], G["psf_a"], 5)
    assert st == 0 and rel_l2(got, G["decon_sv_5it"]) <= 1e-4        # north_star tolerance
    got, st, _ = libapi.decon_dualview(G["img_a"], G["img_b"], G["psf_a"], G["psf_b"], 3)
    assert st == 0 and rel_l2(got, G["decon_dv_3it"]) <= 1e-4
    assert np.array_equal(device.affine_warp(G["img_a"], G["matrix"]), G["warp"])
    r = device.Reg(G["img_a"].shape)
    r.set_images(G["img_a"], G["warp"])
    r.prepare()
    assert np.abs(r.cost(G["cost_matrices"]) - G["costs"]).max() <= 1e-5
    r.close()
    assert device.phasor(G["img_a"], ro.imshift(G["img_a"], (3, -2, 1))) == [3, -2, 1]
