"""Writes tests/golden/path_vectors.npz: small input/output vectors of the hot path as computed by the CPU
oracle (oracle/decon_oracle.py, oracle/reg_oracle.*).  The reference itself ships no vectors and cannot be
built here, so these pin the ORACLE (and through tests/test_gpu_golden.py the CUDA
path) against silent change between rounds; they do not pin it against the reference.
    python tests/golden/make_path_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from microimagelib_b200 import synth  # noqa: E402
from oracle import decon_oracle as do, reg_oracle as ro  # noqa: E402


def inputs():
    psf_a = synth.gaussian_psf((9, 9, 9), (2.0, 1.5, 1.5))
    psf_b = synth.gaussian_psf((9, 9, 9), (1.5, 1.5, 2.0))
    a = synth.bead_image((16, 24, 40), psf_a, density=1 / 256.0, seed=41)
    b = synth.bead_image((16, 24, 40), psf_b, density=1 / 256.0, seed=41, noise_seed=43)
    m = synth.affine_matrix(rot_z_deg=1.0, scale=(1.01, 0.99, 1.0), shift=(0.75, -0.5, 0.25), center=(20, 12, 8))
    return psf_a, psf_b, a, b, m.astype(np.float32)


def compute():
    psf_a, psf_b, a, b, m = inputs()
    out = dict(psf_a=psf_a, psf_b=psf_b, img_a=a, img_b=b, matrix=m)
    out["decon_sv_5it"] = do.decon_singleview(a, psf_a, 5)
    out["decon_dv_3it"] = do.decon_dualview(a, b, psf_a, psf_b, 3)
    src = ro.affine_warp(a, m)
    out["warp"] = src
    t_dm, sd_t = ro.demean(a)
    s_dm, _ = ro.demean(src)
    mats = np.stack([m, np.array([1, 0, 0, 0.5, 0, 1, 0, -0.25, 0, 0, 1, 0], np.float32), np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)])
    out["cost_matrices"] = mats
    out["costs"] = np.array([ro.zncc_cost(t_dm, sd_t, s_dm, k) for k in mats], np.float32)
    out["phasor_shift"] = np.array(ro.phasor(a, ro.imshift(a, (3, -2, 1))), np.int64)
    out["mip_prealign"] = ro.prealign_mip(a, ro.imshift(a, (2.0, -1.0, 1.0)))
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "path_vectors.npz"), **compute())
    print("wrote", os.path.join(HERE, "path_vectors.npz"), os.path.getsize(os.path.join(HERE, "path_vectors.npz")), "bytes")
