"""Synthetic, always-valid bitstreams (no real video is available offline).

ctypes binding of libjsplayer_synth.so (plain C encoders in this directory).  Used by tests/ and bench.py
to make decoder inputs; not part of the decode path.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjsplayer_synth.so")
_lib = None


def build(force=False):
    from . import _build as _b
    return _b.build(force)


class Msv1Recipe(C.Structure):
    _fields_ = [("skip_start_permille", C.c_int32), ("mean_skip", C.c_int32),
                ("pct1", C.c_int32), ("pct2", C.c_int32), ("pct8", C.c_int32)]


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.jsp_synth_msv1_frame.restype = C.c_size_t
        lib.jsp_synth_msv1_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(Msv1Recipe), C.c_void_p, C.c_size_t]
        lib.jsp_synth_msv1_bound.restype = C.c_size_t
        lib.jsp_synth_msv1_bound.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.jsp_sp_enc_new.restype = C.c_void_p
        lib.jsp_sp_enc_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        lib.jsp_sp_enc_free.argtypes = [C.c_void_p]
        lib.jsp_sp_enc_flat.restype = C.c_size_t
        lib.jsp_sp_enc_flat.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]
        lib.jsp_sp_enc_iframe.restype = C.c_size_t
        lib.jsp_sp_enc_iframe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.jsp_sp_enc_pframe.restype = C.c_size_t
        lib.jsp_sp_enc_pframe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        lib.jsp_synth_screen.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p]
        lib.jsp_synth_screen_next.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                              C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.jsp_synth_noise.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.jsp_ans_transitions.argtypes = [C.c_void_p, C.c_int]
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


class SPEncoder:
    """Synthetic ScreenPressor encoder (sp_synth.c). version 2 = range coder, 3/4 = rANS."""

    def __init__(self, width, height, bpp=24, version=2):
        self.lib = load()
        self.w, self.h, self.bpp, self.version = width, height, bpp, version
        self.h_ = self.lib.jsp_sp_enc_new(width, height, bpp, version)
        if not self.h_:
            raise RuntimeError("ScreenPressor encoder for version %d is not available" % version)
        self.cap = width * height * 8 + 4096
        self.buf = np.empty(self.cap, dtype=np.uint8)

    def __del__(self):
        if getattr(self, "h_", None):
            self.lib.jsp_sp_enc_free(self.h_)
            self.h_ = None

    def _px(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        assert a.size == self.w * self.h
        return a

    def iframe(self, px):
        px = self._px(px)
        n = self.lib.jsp_sp_enc_iframe(self.h_, px.ctypes.data, self.buf.ctypes.data, self.cap)
        if n == 0:
            raise ValueError("I frame not representable")
        return self.buf[:n].tobytes()

    def flat(self, colour):
        n = self.lib.jsp_sp_enc_flat(self.h_, int(colour), self.buf.ctypes.data, self.cap)
        return self.buf[:n].tobytes()

    def pframe(self, px, prev, mv=(0, 0)):
        px, prev = self._px(px), self._px(prev)
        n = self.lib.jsp_sp_enc_pframe(self.h_, px.ctypes.data, prev.ctypes.data, int(mv[0]), int(mv[1]), self.buf.ctypes.data, self.cap)
        if n == 0:
            raise ValueError("P frame not representable")
        return self.buf[:n].tobytes()


def screen(width, height, seed, bits=8):
    px = np.empty((height, width), dtype=np.int32)
    load().jsp_synth_screen(width, height, int(seed), bits, px.ctypes.data)
    return px


def screen_next(prev, seed, change_permille=20, bits=8):
    """Returns (picture, (mvx, mvy)) -- the motion vector the scrolled window follows, to offer the encoder."""
    h, w = prev.shape
    px = np.empty((h, w), dtype=np.int32)
    mx, my = C.c_int(0), C.c_int(0)
    load().jsp_synth_screen_next(w, h, int(seed), bits, np.ascontiguousarray(prev).ctypes.data, px.ctypes.data,
                                 int(change_permille), C.byref(mx), C.byref(my))
    return px, (mx.value, my.value)


def sp_stream(width, height, n_frames, seed, version=2, gop=0, change_permille=20, bpp=24):
    """A synthetic stream: frame 0 (and every `gop`-th frame if gop > 0) coded as I frame, the rest as P frames.
    Returns (frames, keys, pictures)."""
    enc = SPEncoder(width, height, bpp, version)
    bits = 8 if bpp != 16 or version != 2 else 5
    frames, keys, pics = [], [], []
    cur = screen(width, height, seed, bits)
    for i in range(n_frames):
        if i == 0 or (gop and i % gop == 0):
            if i:
                cur, _ = screen_next(cur, seed * 7919 + i, change_permille, bits)
            frames.append(enc.iframe(cur)); keys.append(1)
        else:
            nxt, mv = screen_next(cur, seed * 7919 + i, change_permille, bits)
            frames.append(enc.pframe(nxt, cur, mv)); keys.append(0)
            cur = nxt
        pics.append(cur.copy())
    return frames, keys, pics


def noise(width, height, seed, ncolors=(2, 4, 12, 40, 100, 256), bits=8):
    """Noisy picture in vertical bands of ncolors[k]-colour palettes (drives the rANS contexts through all kinds)."""
    px = np.empty((height, width), dtype=np.int32)
    nc = np.asarray(ncolors, dtype=np.int32)
    load().jsp_synth_noise(width, height, int(seed), bits, nc.ctypes.data, int(nc.size), px.ctypes.data)
    return px


ANS_TRANSITIONS = ("4from1", "5from1", "5from4", "6from5", "6from2", "7from3", "7from6", "6grow")


def ans_transitions(reset=False):
    """How often the encoder's rANS colour contexts took each kind transition (coverage of synthetic content)."""
    out = np.zeros(8, dtype=np.int64)
    load().jsp_ans_transitions(out.ctypes.data, int(reset))
    return dict(zip(ANS_TRANSITIONS, (int(v) for v in out)))
