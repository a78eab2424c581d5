"""GPU parity for the caller-side rows of SURVEY.md 8f: the display epilogue (Manager.fill_bitmap_data, reference
src/Manager.hx:363-381, with the render-time flip of src/Main.hx:946) and the key-frame change rule
(Manager.frames_differ_significantly, src/Manager.hx:392-421), both against the oracle's restatement."""
import numpy as np
import pytest

from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def streams():
    out = []
    w, h = 96, 64
    fr = [synth.msv1_frame(False, w, h, 1)] + [synth.msv1_frame(False, w, h, 2 + i, skip_permille=300) for i in range(3)]
    out.append((CodecType.codec_msvc16, O.CODEC_MSVC16, w, h, 16, fr, [1, 0, 0, 0], None))
    pal = synth.random_palette(9)
    fr = [synth.msv1_frame(True, w, h, 7)] + [synth.msv1_frame(True, w, h, 8 + i, skip_permille=200) for i in range(2)]
    out.append((CodecType.codec_msvc8, O.CODEC_MSVC8, w, h, 8, fr, [1, 0, 0], pal))
    fr, k, _ = synth.sp_stream(w, h, 6, seed=3, version=4, gop=3)
    out.append((CodecType.codec_screenpressor, O.CODEC_SCREENPRESSOR, w, h, 24, fr, k, None))
    fr, k, _ = synth.sp_stream(100, 50, 4, seed=4, version=2, bpp=16)          # 16 bpp ScreenPressor: shifted, not swapped
    out.append((CodecType.codec_screenpressor, O.CODEC_SCREENPRESSOR, 100, 50, 16, fr, k, None))
    fr, k, _ = synth.sp_stream(33, 17, 3, seed=5, version=3)                   # width not a multiple of 4: scalar path
    out.append((CodecType.codec_screenpressor, O.CODEC_SCREENPRESSOR, 33, 17, 24, fr, k, None))
    return out


@pytest.mark.parametrize("flip", [False, True])
def test_display_epilogue(flip):
    ss = streams()
    bd = BatchDecoder()
    bd.configure([StreamSpec(c, w, h, bpp, frames=fr, keys=k, palette=pal) for c, _, w, h, bpp, fr, k, pal in ss])
    bd.upload(); bd.run()
    outs, flags = bd.download_display(flip=flip)
    plain, _ = bd.download()
    bd.close()
    i = 0
    for c, oc, w, h, bpp, fr, k, pal in ss:
        exp = O.decode_stream(oc, w, h, bpp, fr, keys=k, palette=pal)[0]
        for f in range(len(fr)):
            assert (plain[i] == exp[f]).all()
            want = O.display_convert(exp[f], from_rgb15=(oc == O.CODEC_SCREENPRESSOR and bpp == 16), flip=flip)
            assert (outs[i] == want).all(), "stream codec %d frame %d" % (oc, f)
            i += 1


def test_key_frame_change_detection():
    w, h = 96, 64
    enc = synth.SPEncoder(w, h, 24, 2)
    a = synth.screen(w, h, 1)
    b, mv = synth.screen_next(a, 2, 200)
    c = a.copy(); c[:20] = b[:20]                                   # differs from `a` only in the first 20 lines
    fa, fb = enc.iframe(a), enc.iframe(b)
    fpb = enc.pframe(b, b)                                          # a P frame that changes nothing visible
    fc = enc.iframe(c)
    fpa = enc.pframe(a, c)
    frames = [fa, fa, fb, fpb, fb, fpa, fc, fpa, enc.iframe(a)]
    keys = [1, 1, 1, 0, 1, 0, 1, 0, 1]
    for insign in (0, 36):
        exp = O.key_frame_differs(O.CODEC_SCREENPRESSOR, w, h, 24, frames, keys, insignificant_lines=insign)
        bd = BatchDecoder(insignificant_lines=insign)
        bd.configure([StreamSpec(CodecType.codec_screenpressor, w, h, 24, frames=frames, keys=keys)])
        outs, flags = bd.decode_host()
        bd.close()
        got = [(int(f) & _lib.JSP_FRAME_DIFFERS) != 0 for f in flags]
        assert got == [bool(x) for x in exp], (insign, got, list(exp))
    # identical consecutive key frames: no change; key after an invisible P frame: pixel compare says no change
    assert list(exp[:5]) == [1, 0, 1, 0, 0]
