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


def _batch_cmd(out_dir, in1, in2, psf_a, psf_b, first, last, reg_mode, initial_tmx="0"):
    # the 34 positional arguments of spimFusionBatch (apps/spim_fusion_batch.cpp: usage)
    return [_need("spimFusionBatch"), str(out_dir) + "/", str(in1) + "/", str(in2) + "/", "A_", "B_", str(first), str(last), "1", str(first),
            "1", "1", "1", "1", "1", "1", str(reg_mode), "0", initial_tmx, "none", "0.001", "200", "0", "1", str(psf_a), str(psf_b), "3",
            "1", "1", "1", "0", "0", "16", "0", "0"]


def _tree(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith((".tif", ".tmx")):
                p = os.path.join(d, f)
                out[os.path.relpath(p, root)] = open(p, "rb").read()
    return out


def test_spim_fusion_batch_pipeline_and_sharding(tmp_path):
    """spimFusionBatch on three small time points: the read-ahead / write-behind I/O pipeline (apps/io_pipeline.h)
    must write exactly the files of the sequential run, and two MILB_SHARD processes together must write them too."""
    from microimagelib_b200 import libapi
    from oracle import reg_oracle as ro
    psf_a = synth.gaussian_psf((13, 13, 13), (2.5, 2, 2))
    psf_b = synth.gaussian_psf((13, 13, 13), (2, 2, 2.5))
    in1, in2 = tmp_path / "SPIMA", tmp_path / "SPIMB"
    in1.mkdir(); in2.mkdir()
    libapi.writetifstack(tmp_path / "pa.tif", psf_a, 32)
    libapi.writetifstack(tmp_path / "pb.tif", psf_b, 32)
    for t in range(3):
        a = synth.bead_image((24, 40, 48), psf_a, density=1 / 512.0, seed=30 + t)
        b = ro.imshift(synth.bead_image((24, 40, 48), psf_b, density=1 / 512.0, seed=30 + t, noise_seed=7), (1, -1, 0))
        libapi.writetifstack(in1 / f"A_{t}.tif", a, 16)
        libapi.writetifstack(in2 / f"B_{t}.tif", b, 16)
    runs = {}
    # seq: host round trips between the stages and no I/O threads (the reference's structure); pipe: I/O threads;
    # resident (the default): the time point stays on the GPU between the stages, 16-bit <-> float on the GPU
    for tag, env in (("seq", {"MILB_PIPELINE": "0", "MILB_DEVICE_RESIDENT": "0"}), ("pipe", {"MILB_PIPELINE": "1", "MILB_DEVICE_RESIDENT": "0"}),
                     ("resident", {"MILB_PIPELINE": "1"}), ("resident_seq", {"MILB_PIPELINE": "0"})):
        out = tmp_path / tag
        r = subprocess.run(_batch_cmd(out, in1, in2, tmp_path / "pa.tif", tmp_path / "pb.tif", 0, 2, 3), capture_output=True, text=True,
                           env={**os.environ, **env}, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        runs[tag] = _tree(out)
    assert len(runs["seq"]) >= 3 * (1 + 1 + 1 + 3)            # per time point: Decon, RegB, matrix, three 2-D MIPs
    for tag in ("pipe", "resident", "resident_seq"):
        assert runs["seq"].keys() == runs[tag].keys()
        for k in runs["seq"]:
            assert runs["seq"][k] == runs[tag][k], (tag, k)
    # two shards (both on GPU 0 here: MILB_SHARD_DEVICE_STRIDE=0) write the same files as the single process
    out = tmp_path / "shards"
    procs = [subprocess.Popen(_batch_cmd(out, in1, in2, tmp_path / "pa.tif", tmp_path / "pb.tif", 0, 2, 3), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env={**os.environ, "MILB_SHARD": f"{r}/2", "MILB_SHARD_DEVICE_STRIDE": "0"})
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs[0][-800:] + outs[1][-800:]
    got = _tree(out)
    assert got.keys() == runs["seq"].keys()
    for k in got:
        assert got[k] == runs["seq"][k], k
