"""Builds the REFERENCE's own libapi (its real GPU path, cuFFT and all) as oracle/_ref/libapi_ref.so.

TEST INFRASTRUCTURE ONLY.  Nothing under microimagelib_b200/, apps/ or bench.py's timed product path
may load this library; it is the checker the oracle and the product are pinned against
(tests/test_gpu_reference_pinned.py, tests/golden/make_reference_golden.py) and the in-run
"reference launch structure on the same GPU" yardstick of bench.py.

The reference (eguomin/microImageLib) does not compile with CUDA 12: include/cukernel.cuh:29-35
declares legacy texture *references*, removed from the toolkit.  Its CPU path needs FFTW3f, which this
image does not have.  This recipe therefore

  1. copies the reference's sources from /root/reference into the git-ignored oracle/_ref/src/
     (they never enter the repository's history),
  2. applies a purely mechanical patch:
       - texture references            -> __device__ cudaTextureObject_t of the same names
         (cukernel.cuh:29,31,35), fetches tex3D(tex, ..) -> tex3D<float>(tex, ..) etc.
       - Bind*/Unbind* host functions  -> oracle/ref_tex_shim.cuh (same filter / address modes,
         including the reference's quirk that BindTexture2 / BindTexture16 configure `tex`, so
         tex2 / tex16 stay at the default point filter)  (src/api_subfunc.cu:885-934, 990-1005)
       - cudaThreadSynchronize         -> cudaDeviceSynchronize
       - lib/tiffconf.h (Windows build of libtiff 4.0.6): __int64 -> long
     No arithmetic, kernel body, launch shape or control flow is touched.
  3. supplies link stubs for what the GPU path never calls: fftwf_* (the gpuMemMode 0 host path;
     a call aborts) -- oracle/ref_link_stubs.c; libtiff comes from the copy bundled with Pillow,
  4. builds with nvcc for sm_100a against cuFFT.

Run here (the reference tree is only present in the build container):  python oracle/build_ref_gpu.py
The resulting .so travels to the GPU box with the snapshot (oracle/_ref/ is git-ignored, not
gpurun-ignored).
"""
from __future__ import annotations

import glob
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MILB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(OUT, "src")
LIB = os.path.join(OUT, "libapi_ref.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _sub(text, pattern, repl, count_min=1, flags=0):
    new, n = re.subn(pattern, repl, text, flags=flags)
    if n < count_min:
        raise RuntimeError(f"patch pattern matched {n} < {count_min} times: {pattern!r}")
    return new


def _remove_function(text, signature_regex):
    """Deletes one brace-balanced function definition whose header matches signature_regex."""
    m = re.search(signature_regex, text)
    if not m:
        raise RuntimeError(f"function not found: {signature_regex!r}")
    i = text.index("{", m.start())
    depth, j = 0, i
    while True:
        c = text[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return text[:m.start()] + text[j + 1:]


def patch_sources():
    os.makedirs(SRC, exist_ok=True)
    for f in ("api_subfunc.cu", "api_decon.cpp", "api_reg.cpp", "api_powell.c", "apifunc.cpp"):
        shutil.copy(os.path.join(REF, "src", f), os.path.join(SRC, f))
    for f in ("cukernel.cuh", "apifunc_internal.h", "powell.h", "libapi.h"):
        shutil.copy(os.path.join(REF, "include", f), os.path.join(SRC, f))
    for f in ("fftw3.h", "tiff.h", "tiffio.h", "tiffconf.h", "tiffvers.h"):
        shutil.copy(os.path.join(REF, "lib", f), os.path.join(SRC, f))
    for f in os.listdir(SRC):
        os.chmod(os.path.join(SRC, f), 0o644)

    # --- cukernel.cuh: texture references -> texture objects -------------------------------------
    p = os.path.join(SRC, "cukernel.cuh")
    t = open(p, encoding="latin-1").read()
    t = _sub(t, r"texture<float, 3, cudaReadModeElementType> tex, tex2;", "__device__ cudaTextureObject_t tex, tex2;")
    t = _sub(t, r"texture<unsigned short, 3, cudaReadModeElementType> tex16;", "__device__ cudaTextureObject_t tex16;")
    t = _sub(t, r"texture<float, 2, cudaReadModeElementType> tex2D1;", "__device__ cudaTextureObject_t tex2D1;")
    t = _sub(t, r"tex3D\(tex16,", "tex3D<unsigned short>(tex16,")
    t = _sub(t, r"tex3D\(tex,", "tex3D<float>(tex,", count_min=3)
    t = _sub(t, r"tex2D\(tex2D1,", "tex2D<float>(tex2D1,", count_min=2)
    open(p, "w", encoding="latin-1").write(t)

    # --- api_subfunc.cu: binding functions -> shim, deprecated sync call ------------------------------
    p = os.path.join(SRC, "api_subfunc.cu")
    t = open(p, encoding="latin-1").read()
    for name in ("BindTexture", "BindTexture2", "BindTexture16", "UnbindTexture", "UnbindTexture2", "UnbindTexture16", "BindTexture2D",
                 "UnbindTexture2D"):
        t = _remove_function(t, r'extern "C" void ' + name + r"\s*\(")
    t = _sub(t, r'#include "cukernel.cuh"', '#include "cukernel.cuh"\n#include "ref_tex_shim.cuh"')
    t = _sub(t, r"cudaThreadSynchronize", "cudaDeviceSynchronize", count_min=10)
    open(p, "w", encoding="latin-1").write(t)
    shutil.copy(os.path.join(HERE, "ref_tex_shim.cuh"), os.path.join(SRC, "ref_tex_shim.cuh"))

    # --- tiffconf.h: the reference ships the Windows configuration of libtiff 4.0.6 (__int64) --------
    p = os.path.join(SRC, "tiffconf.h")
    t = open(p, encoding="latin-1").read()
    t = _sub(t, r"unsigned __int64", "unsigned long")
    t = _sub(t, r"signed __int64", "signed long")
    open(p, "w", encoding="latin-1").write(t)


def find_libtiff():
    import PIL
    libs = glob.glob(os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs", "libtiff-*.so*"))
    return libs[0] if libs else None


def build(verbose=True):
    if not os.path.isdir(os.path.join(REF, "src")):
        if verbose:
            print(f"reference tree absent ({REF}): keeping prebuilt {LIB} (if any)")
        return LIB if os.path.exists(LIB) else None
    patch_sources()
    objs = []
    common = ["-O2", "-w", "-I", SRC, "-Xcompiler", "-fPIC"]
    for f, extra in (("api_subfunc.cu", ["-gencode", "arch=compute_100a,code=sm_100a", "-ftz=true", "-std=c++14"]),
                     ("api_decon.cpp", ["-std=c++14"]), ("api_reg.cpp", ["-std=c++14"]), ("apifunc.cpp", ["-std=c++14"]),
                     ("api_powell.c", [])):
        o = os.path.join(OUT, f + ".o")
        r = subprocess.run([NVCC, *common, *extra, "-c", os.path.join(SRC, f), "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference build failed for {f}:\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}")
        objs.append(o)
    stubs = os.path.join(OUT, "ref_link_stubs.o")
    r = subprocess.run(["gcc", "-O1", "-fPIC", "-c", os.path.join(HERE, "ref_link_stubs.c"), "-o", stubs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("stub build failed:\n" + r.stderr)
    tiff = find_libtiff()
    link = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, stubs, "-lcufft", "-Xlinker", "-Bsymbolic"]
    if tiff:
        link += ["-Xlinker", tiff, "-Xlinker", "--disable-new-dtags", "-Xlinker", "-rpath," + os.path.dirname(tiff)]  # DT_RPATH: libtiff's own dependencies live there too
    else:
        raise RuntimeError("no libtiff found (Pillow's bundled copy expected)")
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference link failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    for o in objs + [stubs]:
        os.remove(o)
    if verbose:
        print(f"built {LIB} from {REF} (patched copy under {SRC})")
    return LIB


if __name__ == "__main__":
    build()
    sys.exit(0)
