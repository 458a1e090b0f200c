"""Builds microimagelib_b200/lib/libapi.so (sm_100a only) with nvcc, in tree.

    python -m microimagelib_b200.build [--force]

nvcc cross-compiles without a GPU.  Objects go to microimagelib_b200/build/ (git-ignored); the
shared library exports the reference's libapi.h entry points plus the milb_* C-ABI
(include/milb_capi.h).  Host code is compiled with -ffp-contract=off: the optimiser and the
parameter->matrix maps must round exactly like the reference's plain C.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# MILB_BUILD_TAG=x builds a variant (with MILB_NVCC_FLAGS) into build_x/ and lib/libapi_x.so for A/B runs (MILB_LIBAPI selects it)
_TAG = os.environ.get("MILB_BUILD_TAG", "")
BUILD = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libapi" + ("_" + _TAG if _TAG else "") + ".so")

SOURCES = ["decon.cu", "decon_fast.cu", "decon_fast_n64.cu", "decon_fast_n128.cu", "decon_fast_n256.cu", "decon_fast_n512.cu", "decon_fast_n1024.cu", "decon_fast_n192.cu", "decon_fast_n320.cu", "decon_fast_n384.cu", "decon_fast_n448.cu", "decon_fast_n576.cu", "decon_fast_n640.cu", "decon_fast_n768.cu", "dslab.cu", "reg.cu", "prealign.cu", "geom.cu", "yardstick.cu", "reg_driver.cpp", "powell.cpp", "libapi.cpp", "tiff_io.cpp"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
EXTRA = os.environ.get("MILB_NVCC_FLAGS", "").split()
COMMON = [*EXTRA, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default", "-DPROJECT_EXPORTS"]


def _deps():
    out = []
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh", ".cu", ".cpp")):
                out.append(os.path.join(root, f))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(BUILD, src + ".o")
    cmd = [NVCC, *ARCH, *COMMON, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _stale(LIB, _deps()):
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"{NVCC} not found and {LIB} is missing or stale")
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcufft"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
