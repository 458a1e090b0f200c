"""apps/io_pipeline.h (the read-ahead / write-behind threads of spimFusionBatch) on the host only: files written
through the worker threads read back exactly, prefetched pairs are handed over, a wrong prefetch is ignored."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "microimagelib_b200", "lib")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(os.path.join(LIB, "libapi.so")):
        pytest.skip("libapi.so not built")
    out = tmp_path_factory.mktemp("iop") / "io_pipeline_check"
    src = os.path.join(ROOT, "tests", "apps", "io_pipeline_check.cpp")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(out), src, "-L" + LIB, "-lapi", "-Wl,-rpath," + LIB,
                        "-Wl,-rpath-link,/usr/local/cuda/lib64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(out)


@pytest.mark.parametrize("pipeline", ["1", "0"])
def test_write_behind_and_read_ahead_round_trip(exe, tmp_path, pipeline):
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env={**os.environ, "MILB_PIPELINE": pipeline}, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("ok")][-1]
    assert f"pipeline={pipeline}" in line
    # pairs 1, 2, 3 and 5, 6 are prefetched; pair 4 was not asked for (pair 0 was), so it is read directly
    assert ("prefetched=5" in line) if pipeline == "1" else ("prefetched=0" in line)
