"""ctypes binding of oracle/liboracle.so -- the CPU restatement of the reference decoders.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing in jsplayer_b200/ imports this module.
"""
import ctypes as C
import fcntl
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
_lib = None

ZERO_STATE, IN_PROGRESS, ERROR_OCCURED = 0, 1, 2
CODEC_SCREENPRESSOR, CODEC_MSVC16, CODEC_MSVC8 = 0, 1, 2


class StreamDesc(C.Structure):
    _fields_ = [("codec", C.c_int), ("width", C.c_int), ("height", C.c_int), ("bpp", C.c_int),
                ("palette", C.c_void_p), ("palette_bytes", C.c_int), ("n_frames", C.c_int),
                ("bytes", C.c_void_p), ("frame_off", C.c_void_p), ("frame_len", C.c_void_p),
                ("frame_key", C.c_void_p), ("out", C.c_void_p)]


def _digest(srcs):
    h = hashlib.sha256()
    for s in sorted(srcs):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False):
    """Builds liboracle.so when its sources changed (content hash, not mtimes: the built file travels with a snapshot of
    the repo); one builder at a time (torchrun ranks)."""
    models = os.path.join(os.path.dirname(HERE), "synth")
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    srcs += [os.path.join(models, f) for f in ("ans_models.c", "ans_models.h")]
    stamp = LIB_PATH + ".srchash"

    def stale():
        try:
            return not os.path.exists(LIB_PATH) or open(stamp).read().strip() != _digest(srcs)
        except OSError:
            return True
    if force or stale():
        with open(LIB_PATH + ".lock", "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or stale():
                    subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True, capture_output=True)
                    with open(stamp, "w") as fh:
                        fh.write(_digest(srcs))
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.ora_create.restype = C.c_void_p
        lib.ora_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.ora_destroy.argtypes = [C.c_void_p]
        lib.ora_preinit.argtypes = [C.c_void_p, C.c_int]
        lib.ora_is_key_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.ora_needs_index.argtypes = [C.c_void_p]
        lib.ora_previous_frame.restype = C.c_void_p
        lib.ora_previous_frame.argtypes = [C.c_void_p]
        lib.ora_decompress_i.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.ora_decompress_p.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        lib.ora_stop_and_clean.argtypes = [C.c_void_p]
        lib.ora_decode_stream.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        lib.ora_display_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.ora_decode_stream_differs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ora_decode_streams_mt.restype = C.c_double
        lib.ora_decode_streams_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        _lib = lib
    return _lib


def _u8(b):
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    return np.ascontiguousarray(a, dtype=np.uint8)


class OracleCodec:
    """IVideoCodec-shaped wrapper (same member names as the reference interface)."""

    def __init__(self, codec, width, height, bpp, palette=None):
        self.lib = load()
        self.X, self.Y = width, height
        pal = _u8(palette) if palette else None
        self._pal = pal
        self.h = self.lib.ora_create(codec, width, height, bpp, pal.ctypes.data if pal is not None else None,
                                     pal.size if pal is not None else 0)
        self._bufs = {}

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ora_destroy(self.h)
            self.h = None

    def Preinit(self, n):
        self.lib.ora_preinit(self.h, n)

    def IsKeyFrame(self, data):
        a = _u8(data)
        return bool(self.lib.ora_is_key_frame(self.h, a.ctypes.data if a.size else None, a.size))

    def NeedsIndex(self):
        return bool(self.lib.ora_needs_index(self.h))

    def PreviousFrame(self):
        return self._bufs.get(self.lib.ora_previous_frame(self.h))

    def DecompressI(self, src, dst):
        a = _u8(src)
        self._bufs[dst.ctypes.data] = dst
        return self.lib.ora_decompress_i(self.h, a.ctypes.data if a.size else None, a.size, dst.ctypes.data)

    def DecompressP(self, src, dst):
        a = _u8(src)
        self._bufs[dst.ctypes.data] = dst
        pnt, sig = C.c_void_p(0), C.c_int(0)
        self.lib.ora_decompress_p(self.h, a.ctypes.data if a.size else None, a.size, dst.ctypes.data, C.byref(pnt), C.byref(sig))
        return self._bufs.get(pnt.value), bool(sig.value)


def decode_stream(codec, width, height, bpp, frames, keys=None, palette=None, insignificant_lines=0):
    """Decodes `frames` (list of bytes) in order; returns (out[n,h,w] int32, changed, significant, status)."""
    lib = load()
    n = len(frames)
    ln = np.array([len(f) for f in frames], dtype=np.uint32)
    off = np.zeros(n, dtype=np.uint64)
    if n:
        off[1:] = np.cumsum(ln.astype(np.uint64))[:-1]
    blob = np.frombuffer(b"".join(bytes(f) for f in frames) + b"\0", dtype=np.uint8).copy()
    k = np.zeros(n, dtype=np.uint8)
    if keys is None:
        if n:
            k[0] = 1
    else:
        k[:] = np.asarray(keys, dtype=np.uint8)
    pal = _u8(palette) if palette else None
    out = np.zeros((n, height, width), dtype=np.int32)
    changed = np.zeros(n, dtype=np.uint8)
    signif = np.zeros(n, dtype=np.uint8)
    status = np.zeros(n, dtype=np.int32)
    lib.ora_decode_stream(codec, width, height, bpp, pal.ctypes.data if pal is not None else None,
                          pal.size if pal is not None else 0, insignificant_lines, n, blob.ctypes.data,
                          off.ctypes.data, ln.ctypes.data, k.ctypes.data, out.ctypes.data, changed.ctypes.data,
                          signif.ctypes.data, status.ctypes.data)
    return out, changed, signif, status


def display_convert(pic, from_rgb15=False, flip=False):
    """Manager.fill_bitmap_data (canvas branch) + optional render-time flip; pic: (h, w) int32."""
    pic = np.ascontiguousarray(pic, dtype=np.int32)
    out = np.empty_like(pic)
    load().ora_display_convert(pic.ctypes.data, out.ctypes.data, pic.shape[1], pic.shape[0], int(from_rgb15), int(flip))
    return out


def key_frame_differs(codec, width, height, bpp, frames, keys, palette=None, insignificant_lines=0):
    """Manager.frames_differ_significantly for every key frame of the stream (0 for the others)."""
    lib = load()
    n = len(frames)
    ln = np.array([len(f) for f in frames], dtype=np.uint32)
    off = np.zeros(n, dtype=np.uint64)
    if n:
        off[1:] = np.cumsum(ln.astype(np.uint64))[:-1]
    blob = np.frombuffer(b"".join(bytes(f) for f in frames) + b"\0", dtype=np.uint8).copy()
    k = np.asarray(keys, dtype=np.uint8).copy()
    pal = _u8(palette) if palette else None
    out = np.zeros((n, height, width), dtype=np.int32)
    differs = np.zeros(n, dtype=np.uint8)
    lib.ora_decode_stream_differs(codec, width, height, bpp, pal.ctypes.data if pal is not None else None,
                                  pal.size if pal is not None else 0, insignificant_lines, n, blob.ctypes.data,
                                  off.ctypes.data, ln.ctypes.data, k.ctypes.data, out.ctypes.data, differs.ctypes.data)
    return differs


def decode_stream_traced(codec, width, height, bpp, frames, keys=None, cap_symbols=4_000_000, **kw):
    """decode_stream plus the oracle's per-symbol trace: an (n_symbols, 5) int32 array of
    (call kind, symbol, freq, cumFreq, total) over the whole stream (see ora_sym_trace in rangecoder_oracle.c).
    Single-threaded test hook."""
    lib = load()
    lib.ora_sym_trace.argtypes = [C.c_void_p, C.c_long]
    lib.ora_sym_trace_count.restype = C.c_long
    buf = np.zeros((cap_symbols, 5), dtype=np.int32)
    lib.ora_sym_trace(buf.ctypes.data, cap_symbols)
    try:
        res = decode_stream(codec, width, height, bpp, frames, keys=keys, **kw)
        n = int(lib.ora_sym_trace_count())
    finally:
        lib.ora_sym_trace(None, 0)
    if n > cap_symbols:
        raise ValueError("trace buffer too small: %d symbols" % n)
    return res + (buf[:n].copy(),)
