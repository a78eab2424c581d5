"""Synthetic, always-valid bitstreams (no real video is available offline).

ctypes binding of libjsplayer_synth.so (plain C encoders in this directory).  Used by tests/ and bench.py
to make decoder inputs; not part of the decode path.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "libjsplayer_synth.so")
_lib = None


class Msv1Recipe(C.Structure):
    _fields_ = [("skip_start_permille", C.c_int32), ("mean_skip", C.c_int32),
                ("pct1", C.c_int32), ("pct2", C.c_int32), ("pct8", C.c_int32)]


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from .. import build
            build.build_synth()
        lib = C.CDLL(LIB_PATH)
        lib.jsp_synth_msv1_frame.restype = C.c_size_t
        lib.jsp_synth_msv1_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(Msv1Recipe), C.c_void_p, C.c_size_t]
        lib.jsp_synth_msv1_bound.restype = C.c_size_t
        lib.jsp_synth_msv1_bound.argtypes = [C.c_int, C.c_int, C.c_int]
        _lib = lib
    return _lib


def msv1_frame(is8, width, height, seed, skip_permille=0, mean_skip=40, mix=(25, 50, 25), out=None):
    """One MSVideo1 frame. skip_permille=0 gives a key frame. Returns bytes, or the length when `out`
    (a uint8 numpy array) is given."""
    lib = load()
    rc = Msv1Recipe(int(skip_permille), int(mean_skip), int(mix[0]), int(mix[1]), int(mix[2]))
    if out is not None:
        n = lib.jsp_synth_msv1_frame(int(bool(is8)), width, height, seed, C.byref(rc), out.ctypes.data, out.size)
        if n == 0 and (width >> 2) * (height >> 2) > 0:
            raise ValueError("output buffer too small")
        return int(n)
    cap = lib.jsp_synth_msv1_bound(int(bool(is8)), width, height)
    buf = np.empty(cap, dtype=np.uint8)
    n = lib.jsp_synth_msv1_frame(int(bool(is8)), width, height, seed, C.byref(rc), buf.ctypes.data, cap)
    return buf[:n].tobytes()


def msv1_bound(is8, width, height):
    return int(load().jsp_synth_msv1_bound(int(bool(is8)), width, height))


def random_palette(seed):
    """256 B,G,R,0 quads as stored after the BITMAPINFOHEADER (strf offset 40)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.integers(0, 256, size=(256, 4), dtype=np.uint8)
    p[:, 3] = 0
    return p.tobytes()
