"""The C-ABI shared library: loads, exports every symbol include/jsplayer_cuda.h declares, and its host-side
entry points (IsKeyFrame, NeedsIndex) agree with the oracle.  No compute calls: runs without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from jsplayer_b200 import _lib
import synth
import jsplayer_b200 as J
from oracle import pyoracle as O


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "jsplayer_cuda.h")).read()
    return sorted(set(re.findall(r"JSP_API[^;(]*?\b(jsp_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    syms = header_symbols()
    assert len(syms) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libjsplayer_cuda.so does not export " + s
    # and the ctypes prototype table covers exactly the header
    assert sorted(_lib.PROTOTYPES) == syms


def test_binding_constants_follow_the_header():
    """Every JSP_* enumerator the ctypes binding restates has the header's value."""
    txt = open(os.path.join(ROOT, "include", "jsplayer_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    vals = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"\b(JSP_[A-Z0-9_]+)\s*=\s*(0x[0-9A-Fa-f]+|\d+)", txt)}
    names = [n for n in dir(_lib) if n.startswith("JSP_") and n in vals]
    assert {"JSP_BATCH_SIGNIFICANCE", "JSP_BATCH_NUMA_BIND", "JSP_BATCH_DISPLAY", "JSP_BATCH_DISPLAY_FLIP", "JSP_DISPLAY_FLIP",
            "JSP_FRAME_CHANGED", "JSP_FRAME_DIFFERS"} <= set(names)
    for n in names:
        assert getattr(_lib, n) == vals[n], n


def test_version_and_error_strings():
    lib = _lib.load()
    assert b"sm_100a" in lib.jsp_version()
    assert lib.jsp_last_error() is not None


@pytest.mark.parametrize("is8", [False, True])
def test_is_key_frame_matches_oracle(is8):
    w, h = 96, 64
    pal = synth.random_palette(3) if is8 else None
    mine = J.MSVideo1_8bit(w, h, pal) if is8 else J.MSVideo1_16bit(w, h)
    ora = O.OracleCodec(O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16, w, h, 8 if is8 else 16, pal)
    assert mine.NeedsIndex() and ora.NeedsIndex()
    rng = np.random.default_rng(5)
    cases = [b"", b"\0", b"\0\0", b"\x04\x84"]
    for s in range(12):
        f = synth.msv1_frame(is8, w, h, 100 + s, skip_permille=0 if s % 2 == 0 else 50)
        cases += [f, f[: len(f) // 2], f[: len(f) // 2 + 1], f + b"\0"]
    cases += [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in (1, 2, 3, 17, 500, 4000)]
    for c in cases:
        assert mine.IsKeyFrame(c) == ora.IsKeyFrame(c)


def test_screenpressor_is_key_frame_heads():
    d = J.ScreenPressor(64, 64, 24)
    assert not d.NeedsIndex()
    for b in range(256):
        assert d.IsKeyFrame(bytes([b, 0, 0])) == (b in (0x11, 0x12, 0x21, 0x22, 0x31, 0x32))   # ScreenPressor.hx:100
    assert not d.IsKeyFrame(b"")


def test_no_gpu_fails_loudly():
    """Without a CUDA device decode calls must fail, never fall back to a CPU path."""
    lib = _lib.load()
    if lib.jsp_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        J.BatchDecoder()
    d = J.MSVideo1_16bit(16, 16)
    dst = np.zeros(256, dtype=np.int32)
    with pytest.raises(RuntimeError):
        d.DecompressI(synth.msv1_frame(False, 16, 16, 1), dst)
    assert lib.jsp_batch_create(0, 0, 0) is None
    assert b"no CUDA device" in lib.jsp_last_error() or b"CUDA" in lib.jsp_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jsplayer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("oracle/msvideo1_oracle.c)", ""), f


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/jsplayer_cuda.h must be consumable by a C compiler (the boundary other hosts bind), and the plain-C
    example host must compile and link against the library.  Without a GPU it exits with the no-fallback message."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    probe = tmp_path / "probe.c"
    probe.write_text('#include "jsplayer_cuda.h"\nint main(void) { return (int)sizeof(jsp_stream_desc) == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, "-c", str(probe), "-o", str(tmp_path / "probe.o")], check=True)
    exe = str(tmp_path / "decode_avi")
    libdir = os.path.join(ROOT, "jsplayer_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", inc, os.path.join(ROOT, "examples", "decode_avi.c"), "-L", libdir,
                    "-ljsplayer_cuda", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if _lib.load().jsp_device_count() <= 0:
        r = subprocess.run([exe, "nonexistent.avi"], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU fallback" in r.stderr
