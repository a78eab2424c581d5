"""Batch driver: many independent streams / GOPs decoded per call (the throughput path).

Host-side counterpart of the frame table the reference's loaders keep per stream
(src/DataLoader.hx:31, src/VideoData.hx:68-73), handed to jsp_batch_* of the C ABI.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .codec import CodecType


@dataclass
class StreamSpec:
    codec: int
    width: int
    height: int
    bpp: int
    frames: Sequence[bytes] = ()                 # compressed frames in order (copied into one buffer), or
    bytes_buf: Optional[np.ndarray] = None        # ... a prebuilt uint8 buffer with frame_off / frame_len
    frame_off: Optional[np.ndarray] = None
    frame_len: Optional[np.ndarray] = None
    keys: Optional[Sequence[int]] = None          # 1 = key frame; default: frame 0 only
    palette: Optional[bytes] = None
    sp_version: int = 0                           # jsp_stream_desc.sp_version: a segment cut out of a longer ScreenPressor stream

    @property
    def n_frames(self):
        return len(self.frame_len) if self.frame_len is not None else len(self.frames)


class PinnedBuffer:
    """cudaHostAlloc'ed memory exposed as a numpy array (jsp_host_alloc / jsp_host_free)."""

    def __init__(self, nbytes, dtype=np.uint8):
        self._lib = _lib.load()
        self.nbytes = int(nbytes)
        self.ptr = self._lib.jsp_host_alloc(max(1, self.nbytes))
        if not self.ptr:
            raise MemoryError("jsp_host_alloc(%d) failed: %s" % (nbytes, _lib.last_error()))
        raw = (C.c_uint8 * max(1, self.nbytes)).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=np.uint8, count=self.nbytes).view(dtype)

    def close(self):
        if self.ptr:
            self.array = None
            self._lib.jsp_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchDecoder:
    def __init__(self, device=-1, insignificant_lines=0, significance=False, numa_bind=False, display=False, display_flip=False):
        """numa_bind: move the calling thread (and the pinned buffers it allocates afterwards) to the device's NUMA node --
        for one-process-per-GPU hosts (JSP_BATCH_NUMA_BIND).
        display: fused display epilogue (JSP_BATCH_DISPLAY, Manager.fill_bitmap_data Manager.hx:363-381) -- the MSVideo1 kernel
        stores canvas R,G,B,A words, with display_flip bottom-up (Main.hx:946); every download then delivers that format."""
        self._lib = _lib.require_gpu()
        self._h = self._lib.jsp_batch_create(int(device), int(insignificant_lines),
                                             (_lib.JSP_BATCH_SIGNIFICANCE if significance else 0) |
                                             (_lib.JSP_BATCH_NUMA_BIND if numa_bind else 0) |
                                             (_lib.JSP_BATCH_DISPLAY if display else 0) |
                                             (_lib.JSP_BATCH_DISPLAY_FLIP if display and display_flip else 0))
        if not self._h:
            raise RuntimeError("jsp_batch_create failed: " + _lib.last_error())
        self._keep = []
        self.specs: List[StreamSpec] = []
        self.n_frames = 0

    def close(self):
        h, self._h = self._h, None
        if h:
            self._lib.jsp_batch_destroy(h)
        self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise RuntimeError("%s failed: %s" % (what, _lib.last_error()))
        return rc

    def configure(self, specs: Sequence[StreamSpec], pinned=False):
        """Builds the C descriptors. With pinned=True frame bytes are gathered into ONE pinned buffer
        (16-byte aligned frames), which is what the end-to-end path wants."""
        self.specs = list(specs)
        keep = []
        descs = (_lib.StreamDescC * len(self.specs))()
        total = 0
        if pinned:
            for sp in self.specs:
                if sp.bytes_buf is None:
                    total += sum((len(f) + 15) & ~15 for f in sp.frames)
            arena = PinnedBuffer(total + 64) if total else None
            keep.append(arena)
            cur = 0
        for i, sp in enumerate(self.specs):
            if sp.bytes_buf is not None:
                buf = sp.bytes_buf
                off = np.ascontiguousarray(sp.frame_off, dtype=np.uint64)
                ln = np.ascontiguousarray(sp.frame_len, dtype=np.uint32)
            else:
                n = len(sp.frames)
                ln = np.array([len(f) for f in sp.frames], dtype=np.uint32)
                al = (ln.astype(np.uint64) + 15) & ~np.uint64(15)
                off = np.zeros(n, dtype=np.uint64)
                if n:
                    off[1:] = np.cumsum(al)[:-1]
                size = int(al.sum())
                if pinned:
                    buf = arena.array[cur:cur + size]
                    cur += size
                else:
                    buf = np.zeros(size + 64, dtype=np.uint8)
                for f, o in zip(sp.frames, off):
                    if len(f):
                        buf[int(o):int(o) + len(f)] = np.frombuffer(f, dtype=np.uint8)
            keys = np.zeros(len(ln), dtype=np.uint8)
            if sp.keys is None:
                if len(keys):
                    keys[0] = 1
            else:
                keys[:] = np.asarray(sp.keys, dtype=np.uint8)
            pal = np.frombuffer(sp.palette, dtype=np.uint8).copy() if sp.palette else None
            keep += [buf, off, ln, keys, pal]
            d = descs[i]
            d.codec, d.width, d.height, d.bpp = int(sp.codec), int(sp.width), int(sp.height), int(sp.bpp)
            d.palette = pal.ctypes.data if pal is not None else None
            d.palette_bytes = int(pal.size) if pal is not None else 0
            d.n_frames = len(ln)
            d.bytes = buf.ctypes.data if buf.size else None
            d.frame_off, d.frame_len, d.frame_key = off.ctypes.data, ln.ctypes.data, keys.ctypes.data
            d.sp_version = int(sp.sp_version)
        self._keep = keep + [descs]
        self.n_frames = int(self._check(self._lib.jsp_batch_configure(self._h, descs, len(self.specs)), "jsp_batch_configure"))
        return self.n_frames

    def frame_shapes(self):
        out = []
        for sp in self.specs:
            out += [(sp.height, sp.width)] * sp.n_frames
        return out

    def upload(self):
        self._check(self._lib.jsp_batch_upload(self._h), "jsp_batch_upload")

    def run(self):
        self._check(self._lib.jsp_batch_run(self._h), "jsp_batch_run")

    def sync(self):
        self._check(self._lib.jsp_batch_sync(self._h), "jsp_batch_sync")

    def _out_ptrs(self, outs):
        ptrs = (C.c_void_p * self.n_frames)()
        for i, a in enumerate(outs):
            if a is not None:
                assert a.dtype == np.int32 and a.flags.c_contiguous
                ptrs[i] = a.ctypes.data
        return ptrs

    def alloc_outputs(self, pinned=False, ring_bytes=0, keep_streams=()):
        """Host pictures for download / decode_host.  ring_bytes > 0 (pinned only): a bounded ring instead of one buffer per
        frame -- what a player does (Manager.hx:424-443 recycles 9 buffers): frame i lands in ring slot i mod R, so every byte
        still crosses PCIe but the host keeps only the last R pictures; the frames of `keep_streams` get buffers of their own
        (so they can be checked afterwards)."""
        shapes = self.frame_shapes()
        if pinned and ring_bytes:
            first = np.cumsum([0] + [sp.n_frames for sp in self.specs])
            own = set()
            for s in keep_streams:
                own.update(range(int(first[s]), int(first[s + 1])))
            biggest = max(h * w for h, w in shapes)
            slots = max(2, int(ring_bytes) // (biggest * 4))
            pb = PinnedBuffer((slots + len(own)) * biggest * 4, np.int32)
            self._keep.append(pb)
            outs, k, r = [], slots, 0
            for i, (h, w) in enumerate(shapes):
                if i in own:
                    outs.append(pb.array[k * biggest:k * biggest + h * w].reshape(h, w)); k += 1
                else:
                    outs.append(pb.array[r * biggest:r * biggest + h * w].reshape(h, w)); r = (r + 1) % slots
            return outs
        if pinned:
            total = sum(h * w for h, w in shapes)
            pb = PinnedBuffer(total * 4, np.int32)
            self._keep.append(pb)
            outs, cur = [], 0
            for h, w in shapes:
                outs.append(pb.array[cur:cur + h * w].reshape(h, w))
                cur += h * w
            return outs
        return [np.zeros((h, w), dtype=np.int32) for h, w in shapes]

    def download(self, outs=None):
        if outs is None:
            outs = self.alloc_outputs()
        flags = np.zeros(self.n_frames, dtype=np.uint8)
        self._check(self._lib.jsp_batch_download(self._h, self._out_ptrs(outs), flags.ctypes.data), "jsp_batch_download")
        return outs, flags

    def download_display(self, outs=None, flip=False):
        """Pictures in the caller's display format (Manager.fill_bitmap_data): canvas R,G,B,A words, optional flip."""
        if outs is None:
            outs = self.alloc_outputs()
        flags = np.zeros(self.n_frames, dtype=np.uint8)
        self._check(self._lib.jsp_batch_download_display(self._h, self._out_ptrs(outs), flags.ctypes.data,
                                                         _lib.JSP_DISPLAY_FLIP if flip else 0), "jsp_batch_download_display")
        return outs, flags

    def results(self):
        flags = np.zeros(self.n_frames, dtype=np.uint8)
        self._check(self._lib.jsp_batch_results(self._h, flags.ctypes.data), "jsp_batch_results")
        return flags

    def next_significant(self, stream, from_frame):
        """Manager.SkipStills' search over the decoded batch: the next frame of `stream` at or after `from_frame` whose change
        is significant (its last frame when there is none)."""
        return int(self._check(self._lib.jsp_batch_next_significant(self._h, int(stream), int(from_frame)), "jsp_batch_next_significant"))

    def decode_host(self, outs=None):
        """upload + decode + download through host buffers (the end-to-end call)."""
        if outs is None:
            outs = self.alloc_outputs()
        flags = np.zeros(self.n_frames, dtype=np.uint8)
        self._check(self._lib.jsp_batch_decode_host(self._h, self._out_ptrs(outs), flags.ctypes.data), "jsp_batch_decode_host")
        return outs, flags

    def alloc_stream_pictures(self, pinned=False):
        """One picture per STREAM (for decode_host_delta)."""
        shapes = [(sp.height, sp.width) for sp in self.specs]
        if pinned:
            pb = PinnedBuffer(sum(h * w for h, w in shapes) * 4, np.int32)
            self._keep.append(pb)
            out, cur = [], 0
            for h, w in shapes:
                out.append(pb.array[cur:cur + h * w].reshape(h, w)); cur += h * w
            return out
        return [np.zeros(s, dtype=np.int32) for s in shapes]

    def decode_host_delta(self, pictures=None, on_frame=None):
        """Opt-in end-to-end path (jsp_batch_decode_host_delta): one picture per stream, updated in place frame by frame from
        the 16x16 blocks that changed; on_frame(stream, frame, picture, flags) sees every frame in order.  Returns
        (pictures, flags): the pictures hold every stream's LAST frame."""
        if pictures is None:
            pictures = self.alloc_stream_pictures()
        ptrs = (C.c_void_p * len(self.specs))()
        for i, a in enumerate(pictures):
            if a is not None:
                assert a.dtype == np.int32 and a.flags.c_contiguous
                ptrs[i] = a.ctypes.data
        flags = np.zeros(self.n_frames, dtype=np.uint8)
        cb = None
        if on_frame is not None:
            def _cb(user, stream, frame, pic, fl):
                on_frame(int(stream), int(frame), pictures[stream], int(fl))
            cb = _lib.FRAME_FN(_cb)
        self._check(self._lib.jsp_batch_decode_host_delta(self._h, ptrs, flags.ctypes.data, cb, None), "jsp_batch_decode_host_delta")
        return pictures, flags

    def delta_bytes(self):
        """Bytes the last decode_host_delta moved device -> host."""
        return int(self._lib.jsp_batch_delta_bytes(self._h))

    def decode(self, specs, significance=None):
        self.configure(specs)
        return self.decode_host()

    def time_runs(self, warmup=3, iters=10, flush_l2=True):
        ms = C.c_float(0)
        kms = (C.c_float * _lib.JSP_N_KERNELS)()
        cnt = (C.c_int64 * _lib.JSP_N_KERNELS)()
        self._check(self._lib.jsp_batch_time_runs(self._h, warmup, iters, 1 if flush_l2 else 0, C.byref(ms), kms, cnt), "jsp_batch_time_runs")
        return ms.value, list(kms), list(cnt)

    def stats(self):
        v = [C.c_uint64(0) for _ in range(4)]
        self._lib.jsp_batch_stats(self._h, *[C.byref(x) for x in v])
        return dict(pixels=v[0].value, alg_bytes=v[1].value, in_bytes=v[2].value, out_bytes=v[3].value)

    def symbols(self):
        """Entropy-coded symbols the ScreenPressor kernels decoded in the last run (0 for MSVideo1-only batches)."""
        return int(self._check(self._lib.jsp_batch_symbols(self._h), "jsp_batch_symbols"))

    def kernel_bytes(self):
        """Algorithmic bytes per run of every kernel class (index = JSP_K_*)."""
        v = (C.c_uint64 * _lib.JSP_N_KERNELS)()
        self._check(self._lib.jsp_batch_kernel_bytes(self._h, v), "jsp_batch_kernel_bytes")
        return [int(x) for x in v]

    def device_frame_ptr(self, i):
        return int(self._lib.jsp_batch_device_frame(self._h, int(i)))
