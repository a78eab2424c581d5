"""AVI files -> batch decoder input: the local-file counterpart of the reference's AVIParser +
DataLoaderAVIIndexed (src/AVIParser.hx:42-184, src/DataLoaderAVIIndexed.hx:276-350, src/DataLoader.hx:321-401).

The RIFF walk and index handling are native (csrc/avi_index.cpp behind jsp_avi_* of the C ABI); this module loads
files into pinned host buffers, resolves key flags the way the reference's loaders do (index first, the codec's
IsKeyFrame otherwise -- DataLoaderAVIIndexed.hx:182) and cuts streams into keyframe-delimited segments (GOPs), the
unit the reference restarts from when seeking (src/Manager.hx:244-249) and the unit of multi-GPU sharding.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .batch import PinnedBuffer, StreamSpec


@dataclass
class AviStream:
    path: Optional[str]
    codec: int
    width: int
    height: int
    bpp: int
    fourcc: int
    fps: float
    has_index: bool
    palette: Optional[bytes]
    data: np.ndarray              # the whole file (uint8; pinned when loaded with pinned=True)
    frame_off: np.ndarray         # uint64 payload offsets into data
    frame_len: np.ndarray         # uint32
    keys: np.ndarray              # uint8
    _pin: object = None

    @property
    def n_frames(self):
        return int(self.frame_len.size)

    def frame(self, i):
        o, n = int(self.frame_off[i]), int(self.frame_len[i])
        return self.data[o:o + n]

    def spec(self, lo=0, hi=None, sp_version=0):
        """StreamSpec for frames [lo, hi) -- no copy: the descriptor points into the file buffer."""
        hi = self.n_frames if hi is None else hi
        return StreamSpec(self.codec, self.width, self.height, self.bpp, bytes_buf=self.data,
                          frame_off=self.frame_off[lo:hi].copy(), frame_len=self.frame_len[lo:hi].copy(),
                          keys=self.keys[lo:hi].copy(), palette=self.palette, sp_version=sp_version)

    def segments(self):
        """[(lo, hi, sp_version)] keyframe-delimited segments (jsp_segment_stream of the C ABI); frames before the first
        key frame form a segment of their own.  sp_version carries the one piece of state ScreenPressor keeps across key
        frames (its entropy coder, ScreenPressor.hx:160-162) into segments cut out of the stream."""
        n = self.n_frames
        if n == 0:
            return []
        first = np.zeros(n, dtype=np.int32)
        ver = np.zeros(n, dtype=np.int32)
        off = np.ascontiguousarray(self.frame_off, dtype=np.uint64)
        ln = np.ascontiguousarray(self.frame_len, dtype=np.uint32)
        keys = np.ascontiguousarray(self.keys, dtype=np.uint8)
        k = _lib.load().jsp_segment_stream(int(self.codec), self.data.ctypes.data, off.ctypes.data, ln.ctypes.data,
                                           keys.ctypes.data, n, first.ctypes.data, ver.ctypes.data)
        if k < 0:
            raise ValueError("jsp_segment_stream failed")
        lo = [int(v) for v in first[:k]]
        return [(a, b, int(v)) for a, b, v in zip(lo, lo[1:] + [n], ver[:k])]

    def gops(self):
        """[(lo, hi)] of segments()."""
        return [(lo, hi) for lo, hi, _ in self.segments()]


def parse_avi(data, path=None, pin=None) -> AviStream:
    """data: uint8 numpy array holding a complete AVI file."""
    lib = _lib.load()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    h = lib.jsp_avi_parse(data.ctypes.data, data.size)
    if not h:
        raise ValueError("%s: %s" % (path or "<memory>", lib.jsp_avi_last_error().decode()))
    try:
        info = _lib.AviInfoC()
        lib.jsp_avi_get_info(h, C.byref(info))
        n = info.n_frames
        off = np.zeros(n, dtype=np.uint64); ln = np.zeros(n, dtype=np.uint32)
        key = np.zeros(n, dtype=np.uint8); known = np.zeros(n, dtype=np.uint8)
        lib.jsp_avi_frame_table(h, off.ctypes.data, ln.ctypes.data, key.ctypes.data, known.ctypes.data)
        pal = None
        if info.palette_bytes > 0:
            buf = np.zeros(info.palette_bytes, dtype=np.uint8)
            lib.jsp_avi_get_palette(h, buf.ctypes.data, buf.size)
            pal = buf.tobytes()
    finally:
        lib.jsp_avi_free(h)
    if not known.all():                        # no index entry: ask the codec (host-side parse, no GPU)
        d = lib.jsp_create(info.codec, info.width, info.height, info.bpp, None, 0, -1)
        try:
            for i in np.nonzero(known == 0)[0]:
                o, m = int(off[i]), int(ln[i])
                key[i] = 1 if m and lib.jsp_is_key_frame(d, data.ctypes.data + o, m) else 0
        finally:
            lib.jsp_destroy(d)
    return AviStream(path, info.codec, info.width, info.height, info.bpp, info.fourcc, info.fps, bool(info.has_index),
                     pal, data, off, ln, key, pin)


def load_avi(path, pinned=False) -> AviStream:
    size = os.path.getsize(path)
    if pinned:
        pb = PinnedBuffer(size + 64)
        data = pb.array[:size]
        with open(path, "rb") as fh:
            fh.readinto(memoryview(data))
        return parse_avi(data, path, pb)
    return parse_avi(np.fromfile(path, dtype=np.uint8), path)


def gop_specs(streams: Sequence[AviStream]):
    """One StreamSpec per keyframe-delimited segment of every file, plus (file, lo, hi) for each."""
    specs, where = [], []
    for fi, st in enumerate(streams):
        for lo, hi, ver in st.segments():
            specs.append(st.spec(lo, hi, ver)); where.append((fi, lo, hi))
    return specs, where


def shard(weights: Sequence[int], n_shards: int) -> List[List[int]]:
    """Longest-first assignment of independent units (streams / GOPs) to n_shards devices or ranks:
    deterministic, so every rank computes the same partition and takes its own part -- no exchange step."""
    order = sorted(range(len(weights)), key=lambda i: (-int(weights[i]), i))
    load = [0] * n_shards
    out = [[] for _ in range(n_shards)]
    for i in order:
        g = min(range(n_shards), key=lambda k: (load[k], k))
        out[g].append(i); load[g] += int(weights[i])
    for part in out:
        part.sort()
    return out
