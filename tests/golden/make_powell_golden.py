"""Writes tests/golden/powell_trajectories.json by running the REFERENCE optimiser
(/root/reference/src/api_powell.c compiled unchanged into oracle/_ref/libpowell_ref.so) on the
analytic float32 cost functions of tests/test_powell_parity.py.  Run in the dev container:
    python tests/golden/make_powell_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import test_powell_parity as tp  # noqa: E402

out = {}
for name, n, ftol, limit in tp.CASES:
    r = tp._run("ref", name, n, ftol, limit)
    h = hashlib.sha256()
    for x, v in r["trace"]:
        h.update(x.tobytes()); h.update(np.float32(v).tobytes())
    out[f"{name}-{n}-{ftol}-{limit}"] = dict(
        n_eval=len(r["trace"]), iter=int(r["iter"]), fret_bits=int(np.float32(r["fret"]).view(np.uint32)),
        p_bits=[int(v) for v in r["p"][1:].view(np.uint32)],
        last_x_bits=[int(v) for v in r["trace"][-1][0].view(np.uint32)],
        trace_sha256=h.hexdigest())
with open(os.path.join(HERE, "powell_trajectories.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print("wrote", len(out), "cases")
