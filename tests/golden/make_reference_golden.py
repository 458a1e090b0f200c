"""Writes tests/golden/reference_vectors.npz: small inputs and the outputs of the REFERENCE ITSELF
(oracle/_ref/libapi_ref.so = the reference's own libapi, built by oracle/build_ref_gpu.py and run on a
B200 through gpurun).  Unlike path_vectors.npz (oracle outputs) these pin the oracle and the CUDA path to
what the reference computes:  tests/test_reference_golden.py checks the oracle against them on the CPU
and the product against them on the GPU.

    gpurun -- 'python tests/golden/make_reference_golden.py gpurun_out/reference_vectors.npz'
    cp gpurun_out/reference_vectors.npz tests/golden/          # then commit
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from microimagelib_b200 import synth  # noqa: E402

IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def inputs():
    psf_a = synth.gaussian_psf((9, 9, 9), (2.0, 1.5, 1.5))
    psf_b = synth.gaussian_psf((9, 9, 9), (1.5, 1.5, 2.0))
    psf_even = synth.gaussian_psf((8, 8, 8), (1.5, 1.5, 1.5))
    a = synth.bead_image((16, 24, 40), psf_a, density=1 / 256.0, seed=41)
    b = synth.bead_image((16, 24, 40), psf_b, density=1 / 256.0, seed=41, noise_seed=43)
    c = synth.bead_image((32, 32, 32), psf_a, density=1 / 256.0, seed=44)          # power-of-two box: the fast kernels
    m = synth.affine_matrix(rot_z_deg=1.0, scale=(1.01, 0.99, 1.0), shift=(0.75, -0.5, 0.25), center=(20, 12, 8)).astype(np.float32)
    return dict(psf_a=psf_a, psf_b=psf_b, psf_even=psf_even, img_a=a, img_b=b, img_c=c, matrix=m)


def compute():
    from oracle import ref_gpu
    R = ref_gpu.api()
    g = inputs()
    a, b, c, m = g["img_a"], g["img_b"], g["img_c"], g["matrix"]
    out = dict(g)
    out["decon_sv_5it"], st, _ = R.decon_singleview(a, g["psf_a"], 5)
    out["decon_sv_even_psf_4it"], st, _ = R.decon_singleview(a, g["psf_even"], 4)
    out["decon_sv_constinit_3it"], st, _ = R.decon_singleview(a, g["psf_a"], 3, initialFlag=True)
    out["decon_sv_unmatched_3it"], st, _ = R.decon_singleview(a, g["psf_a"], 3, flagUnmatch=True, psf_bp=g["psf_b"])
    out["decon_sv_pow2_6it"], st, _ = R.decon_singleview(c, g["psf_a"], 6)
    out["decon_dv_3it"], st, _ = R.decon_dualview(a, b, g["psf_a"], g["psf_b"], 3)
    out["warp"], st = R.atrans3dgpu(a, m)
    src = out["warp"]
    mats = np.stack([m, np.array([1, 0, 0, 0.5, 0, 1, 0, -0.25, 0, 0, 1, 0], np.float32), IDENT])
    out["cost_matrices"] = mats
    zn = []
    for k in mats:       # records[1] of an affMethod-5 call with an input matrix = corrfunc at that matrix (api_subfunc.cu:2817-2821, 2881)
        _, _, st, rec = R.reg3d(a, src, regChoice=2, regMethod=5, inputTmx=True, iTmx=k, itLimit=1)
        zn.append(rec[1])
    out["zncc"] = np.array(zn, np.float32)
    reg, tmx, st, rec = R.reg3d(a, src, regChoice=2, regMethod=7, FTOL=1e-4, itLimit=3000)
    out["reg3d_m7_tmx"], out["reg3d_m7_records"] = tmx, rec
    reg, tmx, st, rec = R.reg3d(a, src, regChoice=2, regMethod=2, FTOL=1e-4, itLimit=3000)
    out["reg3d_m2_tmx"], out["reg3d_m2_records"] = tmx, rec
    z, x, y, st = R.mp2dgpu(a)
    out["mip_z"], out["mip_x"], out["mip_y"] = z, x, y
    out["mip3d_y"], st = R.mip3dgpu(a, 2, 4)
    out["rot_plus90"], st = R.imoperation3D(a, 1)
    out["resize"], st = R.imresize3d(a, (24, 24, 40))
    return out


if __name__ == "__main__":
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(dst, **compute())
    print("wrote", dst, os.path.getsize(dst), "bytes")
