"""The command-line apps end to end on small TIFF stacks (GPU): deconSingleView, deconDualView and
reg3D must write what the oracle computes from the same files."""
import os
import subprocess

import numpy as np
import pytest

from microimagelib_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "apps", "bin")


def _need(app):
    p = os.path.join(BIN, app)
    if not os.path.exists(p):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "apps")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return p


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))


def test_decon_single_and_dual_view_apps(tmp_path):
    from microimagelib_b200 import libapi
    from oracle import decon_oracle as do
    psf_a = synth.gaussian_psf((17, 17, 17), (3, 2, 2))
    psf_b = synth.gaussian_psf((17, 17, 17), (2, 2, 3))
    a = synth.bead_image((24, 40, 56), psf_a, density=1 / 1024.0)
    b = synth.bead_image((24, 40, 56), psf_b, density=1 / 1024.0, noise_seed=9)
    for name, arr in (("a", a), ("b", b), ("pa", psf_a), ("pb", psf_b)):
        libapi.writetifstack(tmp_path / f"{name}.tif", arr, 32)
    out = tmp_path / "sv.tif"
    r = subprocess.run([_need("deconSingleView"), "-i", str(tmp_path / "a.tif"), "-fp", str(tmp_path / "pa.tif"), "-o", str(out), "-it", "5", "-bit", "32"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "runStatus: 0" in r.stdout, r.stdout[-800:] + r.stderr[-800:]
    assert rel_l2(libapi.readtifstack(out), do.decon_singleview(a, psf_a, 5)) <= 1e-4
    out = tmp_path / "dv.tif"
    r = subprocess.run([_need("deconDualView"), "-i1", str(tmp_path / "a.tif"), "-i2", str(tmp_path / "b.tif"), "-fp1", str(tmp_path / "pa.tif"),
                        "-fp2", str(tmp_path / "pb.tif"), "-o", str(out), "-it", "4", "-bit", "16"], capture_output=True, text=True)
    assert r.returncode == 0 and "runStatus: 0" in r.stdout, r.stdout[-800:] + r.stderr[-800:]
    ref = do.decon_dualview(a, b, psf_a, psf_b, 4)
    got = libapi.readtifstack(out)
    # 16-bit output: (uint16) truncation of a float volume that is only 1e-4-close to the oracle's
    assert np.abs(got - np.trunc(ref)).max() <= 1.0 and (got != np.trunc(ref)).mean() < 0.01


def test_reg3d_app_writes_matrix_and_image(tmp_path):
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    psf = synth.gaussian_psf((13, 13, 13), (2, 2, 2))
    tgt = synth.bead_image((24, 32, 40), psf, density=1 / 512.0, seed=5)
    m = synth.affine_matrix(rot_z_deg=1.5, scale=(1.01, 0.99, 1.0), shift=(0.8, -0.6, 0.4), center=(20, 16, 12))
    src = synth.warp_exact(tgt, m)
    libapi.writetifstack(tmp_path / "t.tif", tgt, 32)
    libapi.writetifstack(tmp_path / "s.tif", src, 32)
    r = subprocess.run([_need("reg3D"), "-t", str(tmp_path / "t.tif"), "-s", str(tmp_path / "s.tif"), "-o", str(tmp_path / "r.tif"),
                        "-otmx", str(tmp_path / "m.tmx"), "-affm", "6", "-bit", "32", "-verbOFF"], capture_output=True, text=True)
    assert r.returncode == 0 and "runStatus: 0" in r.stdout, r.stdout[-800:] + r.stderr[-800:]
    ref = ro.reg3d_affine(tgt, src, 6)
    rows = [l.split() for l in open(tmp_path / "m.tmx").read().strip().splitlines()]
    assert len(rows) == 4 and rows[3] == ["0.000000", "0.000000", "0.000000", "1.000000"]
    got = np.array([float(v) for row in rows[:3] for v in row], np.float32)
    assert np.abs(got - ref["tmx"]).max() < 1e-5            # "%f" text keeps 6 decimals
    assert np.array_equal(libapi.readtifstack(tmp_path / "r.tif"), ref["reg"])
