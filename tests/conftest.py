import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "hw_fetch: run with the product's default ZNCC source fetch (hardware texture unit)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """The oracle's C restatement (and, where the reference tree exists, oracle/_ref) must be
    built; the prebuilt .so files travel to the GPU box."""
    need = [os.path.join(ROOT, "oracle", "liboracle_reg.so")]
    if not all(os.path.exists(p) for p in need) or os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=False, capture_output=True)
    yield


# The warp, registration-cost and rotating-projection kernels sample the source with the hardware texture unit by default
# (the reference's own mechanism: each sample is the float the reference's tex3D returns).  The CPU oracle restates that
# fetch in software; its bit-identical CUDA twin is selected with MILB_TEX_FETCH=sw.  Tests that compare Powell trajectories / sums
# with the ORACLE bit for bit run the twin; tests marked hw_fetch and tests/test_gpu_reference_pinned.py (product vs
# the reference itself) run the default.
_ORACLE_TWIN_MODULES = ("test_gpu_reg", "test_gpu_prealign", "test_gpu_apps", "test_golden_vectors", "test_gpu_geometry")


@pytest.fixture(autouse=True)
def _zncc_fetch_mode(request, monkeypatch):
    mod = request.module.__name__.split(".")[-1]
    if mod in _ORACLE_TWIN_MODULES and "hw_fetch" not in request.keywords:
        monkeypatch.setenv("MILB_TEX_FETCH", "sw")
    yield
