"""TIFF stack I/O of libapi (csrc/tiff_io.cpp, host only): round trips, the reference's pixel
conversion rules (src/apifunc.cpp:171-175, :255) and interoperability with libtiff via Pillow."""
import os

import numpy as np
import pytest

from microimagelib_b200 import libapi


def test_roundtrip_16_and_32_bit(tmp_path):
    rng = np.random.default_rng(0)
    vol = (rng.random((5, 7, 9)) * 4000).astype(np.float32)
    p16, p32 = tmp_path / "a16.tif", tmp_path / "a32.tif"
    libapi.writetifstack(p16, vol, 16)
    libapi.writetifstack(p32, vol, 32)
    assert libapi.gettifinfo(p16) == (16, (9, 7, 5))
    assert libapi.gettifinfo(p32) == (32, (9, 7, 5))
    assert np.array_equal(libapi.readtifstack(p32), vol)                      # float32: bit exact
    assert np.array_equal(libapi.readtifstack(p16), np.trunc(vol))            # (uint16) truncation


def test_uint16_conversion_truncates_and_wraps_like_x86(tmp_path):
    vol = np.array([[[0.9, 1.5, 65535.7, 65536.0, 70000.2, -1.5, -0.4]]], np.float32)
    p = tmp_path / "w.tif"
    libapi.writetifstack(p, vol, 16)
    got = libapi.readtifstack(p)
    assert got.ravel().tolist() == [0, 1, 65535, 0, 4464, 65535, 0]


def _as_contig_planar(path, out):
    """Pillow cannot decode PLANARCONFIG_SEPARATE (which the reference sets, src/apifunc.cpp:266, and
    which is meaningless for one sample per pixel); flip tag 284 to CONTIG in a copy for the pixel check."""
    import struct
    b = bytearray(open(path, "rb").read())
    key = struct.pack("<HHI", 284, 3, 1)
    i = b.find(key)
    while i >= 0:
        b[i + 8:i + 10] = struct.pack("<H", 1)
        i = b.find(key, i + 12)
    open(out, "wb").write(b)
    return out


def test_pillow_reads_what_we_write_and_vice_versa(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(1)
    vol = (rng.random((4, 6, 8)) * 1000).astype(np.float32)
    p = tmp_path / "ours.tif"
    libapi.writetifstack(p, vol, 16)
    im = Image.open(p)
    assert getattr(im, "n_frames", 1) == 4 and im.mode == "I;16"
    tags = dict(im.tag_v2)
    # the reference's tag set, src/apifunc.cpp:260-271
    assert tags[256] == 8 and tags[257] == 6 and tags[258] == (16,) and tags[259] == 1 and tags[262] == 1
    assert tags[274] == 1 and tags[277] == 1 and tags[278] == 6 and tags[279] == (96,) and tags[284] == 2
    im = Image.open(_as_contig_planar(p, tmp_path / "ours_c.tif"))
    for k in range(4):
        im.seek(k)
        assert np.array_equal(np.array(im), np.trunc(vol[k]).astype(np.uint16))
    p = tmp_path / "ours32.tif"
    libapi.writetifstack(p, vol, 32)
    assert dict(Image.open(p).tag_v2)[339] in (3, (3,))          # SAMPLEFORMAT_IEEEFP
    im = Image.open(_as_contig_planar(p, tmp_path / "ours32_c.tif"))
    im.seek(2)
    assert np.array_equal(np.array(im), vol[2])
    # a stack written by libtiff (through Pillow), uncompressed, possibly several strips per page
    pages = [Image.fromarray((vol[k]).astype(np.uint16)) for k in range(4)]
    q = tmp_path / "theirs.tif"
    pages[0].save(q, save_all=True, append_images=pages[1:], compression=None)
    assert libapi.gettifinfo(q) == (16, (8, 6, 4))
    assert np.array_equal(libapi.readtifstack(q), vol.astype(np.uint16).astype(np.float32))


def _run_reader(path):
    """readtifstack in a child process (the library follows the reference's convention: errors print and exit(1))"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from microimagelib_b200 import libapi\n"
            "v = libapi.readtifstack(%r)\n"
            "print('READ', v.shape)\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(path))
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)


def test_reader_rejects_inconsistent_and_cyclic_stacks(tmp_path):
    """A stack whose pages differ in size, or whose directory chain loops, is refused before anything is copied."""
    from PIL import Image
    from microimagelib_b200 import libapi
    good = tmp_path / "good.tif"
    vol = (np.arange(3 * 6 * 8).reshape(3, 6, 8) % 200).astype(np.float32)
    libapi.writetifstack(good, vol, 16)
    r = _run_reader(good)
    assert r.returncode == 0 and "READ (3, 6, 8)" in r.stdout
    # pages of different sizes (written by Pillow)
    bad = tmp_path / "ragged.tif"
    pages = [Image.fromarray(np.zeros((6, 8), np.uint16)), Image.fromarray(np.zeros((7, 9), np.uint16))]
    pages[0].save(bad, save_all=True, append_images=pages[1:], compression=None)
    r = _run_reader(bad)
    assert r.returncode != 0 and "differ" in (r.stdout + r.stderr)
    # a directory chain that points back to the first directory
    raw = bytearray(open(good, "rb").read())
    first = int.from_bytes(raw[4:8], "little")
    off = first
    while True:
        n = int.from_bytes(raw[off:off + 2], "little")
        nxt_at = off + 2 + 12 * n
        nxt = int.from_bytes(raw[nxt_at:nxt_at + 4], "little")
        if nxt == 0:
            raw[nxt_at:nxt_at + 4] = first.to_bytes(4, "little")
            break
        off = nxt
    cyc = tmp_path / "cyclic.tif"
    open(cyc, "wb").write(bytes(raw))
    r = _run_reader(cyc)
    assert r.returncode != 0
