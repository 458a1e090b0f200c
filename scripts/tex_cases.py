"""Reference warps (hardware tex3D) of a white-noise volume under several affine matrices -> gpurun_out/tex_cases.npz,
for off-line analysis of where the software restatement of the texture filter differs from the hardware (CPU side:
scripts/tex_cases_analyze.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from microimagelib_b200 import synth
from oracle import ref_gpu

def cases():
    rng = np.random.default_rng(77)
    vol = (rng.random((24, 28, 32)) * 1000).astype(np.float32)
    mats = [synth.affine_matrix(rot_z_deg=2.0, scale=(1.02, 0.99, 1.0), shift=(1.5, -1.25, 0.75), center=(16, 14, 12)),
            synth.affine_matrix(rot_z_deg=-3.0, scale=(0.97, 1.03, 1.01), shift=(0.3, 0.7, -0.4), center=(16, 14, 12)),
            synth.affine_matrix(rot_z_deg=0.0, scale=(1.0, 1.0, 1.0), shift=(0.123, 0.456, 0.789)),
            synth.affine_matrix(rot_z_deg=17.0, scale=(1.1, 0.9, 1.05), shift=(2.2, -1.1, 0.6), center=(16, 14, 12)),
            synth.affine_matrix(rot_z_deg=0.0, scale=(1.5, 0.75, 1.25), shift=(0.0, 0.0, 0.0)),
            synth.affine_matrix(rot_z_deg=45.0, scale=(1.0, 1.0, 1.0), shift=(0.5, 0.5, 0.5), center=(16, 14, 12))]
    return vol, np.stack(mats).astype(np.float32)

if __name__ == "__main__":
    vol, mats = cases()
    R = ref_gpu.api()
    outs = np.stack([R.atrans3dgpu(vol, m)[0] for m in mats])
    np.savez_compressed("gpurun_out/tex_cases.npz", mats=mats, outs=outs)
    print("wrote gpurun_out/tex_cases.npz", outs.shape)
