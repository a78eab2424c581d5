"""Host-side mirror of the reference's plugin interface for the codec hot path.

Names, argument meaning and return values follow `interface IVideoCodec`
(reference src/IVideoCodec.hx:16-29) and its three implementations
(src/MSVideo1.hx:8 MSVideo1_16bit, :262 MSVideo1_8bit, src/ScreenPressor.hx:19 ScreenPressor),
so a test written against the reference classes reads the same here.  Frame buffers are
numpy int32 arrays of X*Y elements (the reference's Int32Array), bitstreams are bytes-like
(Uint8Array).  Every Decompress* call runs on the GPU through libjsplayer_cuda's C ABI.
"""
import ctypes as C
import enum
from collections import namedtuple

import numpy as np

from . import _lib


class DecoderState(enum.IntEnum):          # IVideoCodec.hx:5-9
    zero_state = 0
    in_progress = 1
    error_occured = 2


class CodecType(enum.IntEnum):             # VideoData.hx:75-80
    codec_screenpressor = 0
    codec_msvc16 = 1
    codec_msvc8 = 2


# IVideoCodec.hx:11-14; data_pnt is the array object the caller passed (dst, or the retained previous one) or None
PFrameResult = namedtuple("PFrameResult", ["data_pnt", "significant_changes"])


def _u8(data):
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    if a.dtype != np.uint8 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class IVideoCodec:
    """Common implementation over the C ABI; subclasses only choose the codec."""

    _codec = None

    def __init__(self, width, height, bpp, palette=None, device=-1):
        self._lib = _lib.load()
        self.X, self.Y, self.bpp = int(width), int(height), int(bpp)
        pal = _u8(palette) if palette is not None and len(palette) else None
        self._h = self._lib.jsp_create(int(self._codec), self.X, self.Y, self.bpp,
                                       pal.ctypes.data if pal is not None else None,
                                       int(pal.size) if pal is not None else 0, int(device))
        if not self._h:
            raise RuntimeError("jsp_create failed: " + _lib.last_error())
        self._buffers = {}          # address -> numpy array the caller handed in (to return the same object)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.jsp_destroy(h)

    def _dst(self, dst):
        if not (isinstance(dst, np.ndarray) and dst.dtype == np.int32 and dst.flags.c_contiguous
                and dst.size >= self.X * self.Y):
            raise TypeError("dst must be a C-contiguous int32 numpy array with at least X*Y elements")
        self._buffers[dst.ctypes.data] = dst
        return dst.ctypes.data

    def _by_addr(self, addr):
        return self._buffers.get(addr) if addr else None

    # ---- IVideoCodec members ----
    def Preinit(self, insignificant_lines):
        self._lib.jsp_preinit(self._h, int(insignificant_lines))

    def PreviousFrame(self):
        return self._by_addr(self._lib.jsp_previous_frame(self._h))

    def IsKeyFrame(self, data):
        a = _u8(data)
        return bool(self._lib.jsp_is_key_frame(self._h, a.ctypes.data if a.size else None, int(a.size)))

    def State(self):
        return DecoderState(self._lib.jsp_state_of(self._h))

    def DecompressI(self, src, dst):
        a = _u8(src)
        _lib.require_gpu()
        return DecoderState(self._lib.jsp_decompress_i(self._h, a.ctypes.data if a.size else None, int(a.size), self._dst(dst)))

    def ContinueI(self):
        return DecoderState(self._lib.jsp_continue_i(self._h))

    def DecompressP(self, src, dst):
        a = _u8(src)
        _lib.require_gpu()
        r = self._lib.jsp_decompress_p(self._h, a.ctypes.data if a.size else None, int(a.size), self._dst(dst))
        return PFrameResult(self._by_addr(r.data_pnt), bool(r.significant_changes))

    def NeedsIndex(self):
        return bool(self._lib.jsp_needs_index(self._h))

    def StopAndClean(self):
        self._lib.jsp_stop_and_clean(self._h)


class MSVideo1_16bit(IVideoCodec):
    """`new MSVideo1_16bit(width, height)` -- reference src/MSVideo1.hx:20-31."""
    _codec = CodecType.codec_msvc16

    def __init__(self, width, height, device=-1):
        super().__init__(width, height, 16, None, device)


class MSVideo1_8bit(IVideoCodec):
    """`new MSVideo1_8bit(width, height, palette)` -- reference src/MSVideo1.hx:267-274."""
    _codec = CodecType.codec_msvc8

    def __init__(self, width, height, palette, device=-1):
        super().__init__(width, height, 8, palette, device)


class ScreenPressor(IVideoCodec):
    """`new ScreenPressor(width, height, bits_per_pixel)` -- reference src/ScreenPressor.hx:53-64."""
    _codec = CodecType.codec_screenpressor

    def __init__(self, width, height, bits_per_pixel, device=-1):
        super().__init__(width, height, bits_per_pixel, None, device)
