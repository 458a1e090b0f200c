#!/usr/bin/env python
"""bench_fusion.py -- BASELINE's second metric, "fusion vols/sec": spimFusionBatch (registration of a test
time point, then resample + apply matrix + joint dual-view RL deconvolution + 2-D MIPs per time point)
on config-5-like data: T time points of 512x512x256 dual-view uint16 TIFF stacks, INCLUDING disk I/O.

    python bench_fusion.py --points 6 --iters 10 [--gpus N]

Prints one JSON line: time points per second through the app with the read-ahead / write-behind I/O
pipeline on (default) and off (the reference's read -> compute -> write sequence), same files out.
With --gpus N the time points are sharded over N processes (MILB_SHARD=r/N, one GPU each).
Data: synthetic beads, the same pair copied to every time point (timing does not depend on content).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="256,512,512")
    ap.add_argument("--points", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dir", default=None, help="scratch directory (default: a temporary directory)")
    args = ap.parse_args()
    import numpy as np
    from microimagelib_b200 import libapi, synth

    shape = tuple(int(v) for v in args.shape.split(","))
    app = os.path.join(ROOT, "apps", "bin", "spimFusionBatch")
    if not os.path.exists(app):
        subprocess.run(["make", "-C", os.path.join(ROOT, "apps")], check=True, capture_output=True)
    work = args.dir or tempfile.mkdtemp(prefix="milb_fusion_")
    in1, in2 = os.path.join(work, "SPIMA"), os.path.join(work, "SPIMB")
    os.makedirs(in1, exist_ok=True)
    os.makedirs(in2, exist_ok=True)
    psf_a = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
    psf_b = synth.gaussian_psf((65, 65, 65), (2, 2, 4))
    libapi.writetifstack(os.path.join(work, "pa.tif"), psf_a, 32)
    libapi.writetifstack(os.path.join(work, "pb.tif"), psf_b, 32)
    a = synth.bead_image(shape, psf_a, seed=20260)
    b = synth.shift_zero_fill(synth.bead_image(shape, psf_b, seed=20260, noise_seed=20263), (2, -1, 1))
    libapi.writetifstack(os.path.join(in1, "A_0.tif"), a, 16)
    libapi.writetifstack(os.path.join(in2, "B_0.tif"), b, 16)
    for t in range(1, args.points):
        shutil.copyfile(os.path.join(in1, "A_0.tif"), os.path.join(in1, f"A_{t}.tif"))
        shutil.copyfile(os.path.join(in2, "B_0.tif"), os.path.join(in2, f"B_{t}.tif"))
    in_bytes = 2 * os.path.getsize(os.path.join(in1, "A_0.tif"))

    def cmd(out):
        # regMode 1: register the test time point (index 0), then every time point applies that matrix
        return [app, out + "/", in1 + "/", in2 + "/", "A_", "B_", "0", str(args.points - 1), "1", "0", "1", "1", "1", "1", "1", "1", "1", "0", "0",
                "none", "0.001", "1000", "0", "0", os.path.join(work, "pa.tif"), os.path.join(work, "pb.tif"), str(args.iters),
                "1", "1", "1", "0", "0", "16", "0", "0"]

    res = {}
    for tag, pipe in (("pipelined", "1"), ("sequential", "0")):
        out = os.path.join(work, "out_" + tag)
        shutil.rmtree(out, ignore_errors=True)
        t0 = time.perf_counter()
        if args.gpus == 1:
            procs = [subprocess.Popen(cmd(out), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env={**os.environ, "MILB_PIPELINE": pipe})]
        else:
            procs = [subprocess.Popen(cmd(out), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                      env={**os.environ, "MILB_PIPELINE": pipe, "MILB_SHARD": f"{r}/{args.gpus}"}) for r in range(args.gpus)]
        logs = [p.communicate()[0] for p in procs]
        dt = time.perf_counter() - t0
        if any(p.returncode != 0 for p in procs):
            print(logs[0][-2000:], file=sys.stderr)
            raise SystemExit("spimFusionBatch failed")
        n_out = sum(1 for f in os.listdir(os.path.join(out, "Decon")) if f.startswith("Decon_"))
        reg_s = [float(l.split(":")[1].split()[0]) for l in logs[0].splitlines() if l.strip().startswith("Time cost for  registration")]
        stages = [l.strip() for l in logs[0].splitlines() if "Time cost for" in l][-5:]
        per_point = [float(l.split(" is ")[1].split()[0]) for l in logs[0].splitlines() if l.startswith("...Time cost for current image")]
        steady = per_point[2:] if len(per_point) > 3 else per_point
        res[tag] = {"steady_state_s_per_time_point": sum(steady) / max(len(steady), 1), "last_time_point_stages": stages, "wall_s": dt, "vols_per_s": n_out / dt, "volumes_written": n_out, "first_registration_s": reg_s[0] if reg_s else None}
        shutil.rmtree(out, ignore_errors=True)
    line = {"metric": "fusion vols/sec (spimFusionBatch incl. TIFF I/O)", "value": res["pipelined"]["vols_per_s"], "unit": "time points/s",
            "n_gpus": args.gpus, "scaling": "weak" if args.gpus > 1 else None,
            "config": {"workload": f"{args.points} time points, {shape[2]}x{shape[1]}x{shape[0]} dual-view uint16 TIFF pairs ({in_bytes / 1e6:.0f} MB read per time point), "
                                   f"registration mode 1 (test time point, affine 12 DOF), {args.iters} joint RL iterations, X/Y/Z MIPs, 16-bit outputs",
                       "sharding": "MILB_SHARD=r/N, one process and GPU per shard" if args.gpus > 1 else "single process"},
            "steady_state_vols_per_s": args.gpus / res["pipelined"]["steady_state_s_per_time_point"],
            "note": "value = time points / wall time of the whole batch process (start-up, OTF preparation and the test registration included); "
                    "steady_state = 1 / mean per-time-point time after the first two",
            "pipelined": res["pipelined"], "sequential_like_reference": res["sequential"],
            "speedup_from_io_pipeline": res["pipelined"]["vols_per_s"] / res["sequential"]["vols_per_s"], "data": "synthetic"}
    print(json.dumps(line), flush=True)
    if args.dir is None:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
