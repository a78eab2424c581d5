"""The opt-in in-place end-to-end path (jsp_batch_decode_host_delta): only the 16x16 blocks that differ from a stream's
previous picture cross PCIe; the host's one-picture-per-stream buffers must show every frame of every stream, in order,
bit-exact against the oracle."""
import numpy as np
import pytest

from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def streams():
    out = []
    for i, (w, h) in enumerate([(320, 240), (250, 130), (33, 17)]):                      # sizes with partial 16x16 blocks
        frames, keys, _ = synth.sp_stream(w, h, 10, seed=40 + i, version=(2, 3, 4)[i], gop=5, change_permille=50)
        out.append((O.CODEC_SCREENPRESSOR, CodecType.codec_screenpressor, w, h, 24, None, frames, keys))
    for is8, (w, h) in ((False, (320, 240)), (True, (132, 100)), (False, (66, 50))):       # 132, 66, 50: remainders mod 4 and mod 16
        pal = synth.random_palette(5) if is8 else None
        frames = [synth.msv1_frame(is8, w, h, 300)] + [synth.msv1_frame(is8, w, h, 301 + f, skip_permille=800, mean_skip=30) for f in range(7)]
        out.append((O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16, CodecType.codec_msvc8 if is8 else CodecType.codec_msvc16, w, h,
                    8 if is8 else 16, pal, frames, None))
    return out


def test_in_place_delta_path_shows_every_frame():
    ss = streams()
    specs = [StreamSpec(c, w, h, bpp, frames=fr, keys=k, palette=pal) for _, c, w, h, bpp, pal, fr, k in ss]
    exp = [O.decode_stream(oc, w, h, bpp, fr, keys=k, palette=pal, insignificant_lines=36) for oc, _, w, h, bpp, pal, fr, k in ss]
    bd = BatchDecoder(insignificant_lines=36, significance=True)
    bd.configure(specs, pinned=True)
    pics = bd.alloc_stream_pictures()
    for p in pics:
        p[:] = 0x55AA55                                     # junk: the first frame has to overwrite everything the codec writes
    seen = []

    def on_frame(stream, frame, picture, flags):
        oc, _, w, h, bpp, *_ = ss[stream]
        e = exp[stream][0][frame]
        if oc != O.CODEC_SCREENPRESSOR:                      # MSVideo1 never writes the remainder mod 4
            bw, bh = w & ~3, h & ~3
            assert (picture[:bh, :bw] == e[:bh, :bw]).all(), (stream, frame)
            assert (picture[bh:, :] == 0x55AA55).all() and (picture[:, bw:] == 0x55AA55).all()
        else:
            assert (picture == e).all(), (stream, frame)
        assert bool(flags & _lib.JSP_FRAME_CHANGED) == bool(exp[stream][1][frame])
        seen.append((stream, frame))

    pics, flags = bd.decode_host_delta(pics, on_frame)
    for s in range(len(ss)):
        assert [f for st, f in seen if st == s] == list(range(len(ss[s][6])))          # every frame, in order
    # and again without a callback, into fresh buffers: the pictures end on every stream's last frame
    pics2, _ = bd.decode_host_delta()
    bd.close()
    for s, (oc, _, w, h, *_r) in enumerate(ss):
        bw, bh = (w, h) if oc == O.CODEC_SCREENPRESSOR else (w & ~3, h & ~3)
        assert (pics2[s][:bh, :bw] == exp[s][0][-1][:bh, :bw]).all()
