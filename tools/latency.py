"""Per-call latency of the drop-in codec path (one frame per DecompressI / DecompressP call through the C ABI, host
buffers in, host buffers out) next to the CPU oracle decoding the same frames one at a time.  The reference paces its
worker at one frame per 1 ms timer tick (Manager.hx:139-141); this shows what a single stream sees -- the batch path
(bench.py) is where the GPU pays off.  Usage (GPU box): python tools/latency.py"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import MSVideo1_16bit, ScreenPressor, DecoderState
import synth
from jsplayer_b200.batch import PinnedBuffer
from oracle import pyoracle as O


def timed(dec, frames, keys, w, h, pinned):
    bufs = []
    for i in range(3):
        if pinned:
            pb = PinnedBuffer(w * h * 4, np.int32); bufs.append((pb, pb.array.reshape(h, w)))
        else:
            bufs.append((None, np.zeros((h, w), dtype=np.int32)))
    ts = []
    prev = None
    for i, f in enumerate(frames):
        dst = next(b[1] for b in bufs if b[1] is not prev)          # never the retained buffer
        t0 = time.perf_counter()
        if keys[i]:
            st = dec.DecompressI(f, dst); res = dst
        else:
            r = dec.DecompressP(f, dst); res = r.data_pnt
        ts.append(time.perf_counter() - t0)
        if res is not None: prev = res
    return np.array(ts) * 1e3


def oracle_ms(codec, w, h, bpp, frames, keys):
    t0 = time.perf_counter()
    O.decode_stream(codec, w, h, bpp, frames, keys=keys)
    return (time.perf_counter() - t0) * 1e3 / len(frames)


w, h = 1920, 1080
msv = [synth.msv1_frame(False, w, h, 1)] + [synth.msv1_frame(False, w, h, 2 + i, skip_permille=700, mean_skip=30) for i in range(31)]
mkeys = [1] + [0] * 31
sp, skeys, _ = synth.sp_stream(w, h, 32, seed=5, version=2, gop=0, change_permille=20)
sp4, skeys4, _ = synth.sp_stream(w, h, 32, seed=5, version=4, gop=0, change_permille=20)
for name, mk, frames, keys, codec, bpp in (("MSVideo1 RGB555 1080p", lambda: MSVideo1_16bit(w, h), msv, mkeys, O.CODEC_MSVC16, 16),
                                           ("ScreenPressor v2 1080p", lambda: ScreenPressor(w, h, 24), sp, skeys, O.CODEC_SCREENPRESSOR, 24),
                                           ("ScreenPressor v4 1080p", lambda: ScreenPressor(w, h, 24), sp4, skeys4, O.CODEC_SCREENPRESSOR, 24)):
    for pinned in (False, True):
        dec = mk(); dec.Preinit(36)
        timed(dec, frames, keys, w, h, pinned)                      # warm-up pass (device allocations, first-touch)
        t = timed(dec, frames, keys, w, h, pinned)                  # the same stream again, from its key frame
        dec.StopAndClean()
        print("%-24s %-8s key frame %7.2f ms   inter frames median %6.2f ms  max %6.2f ms" % (name, "pinned" if pinned else "pageable", t[0], np.median(t[1:]), t[1:].max()))
    print("%-24s CPU oracle, 1 thread: %6.2f ms per frame (stream average)" % (name, oracle_ms(codec, w, h, bpp, frames, keys)))
