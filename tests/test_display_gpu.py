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


def _msv1_cases():
    """MSVideo1 streams that reach every store of the decode kernel: coded blocks of the three classes, short skip runs,
    pre-copied sparse frames, "rest of the frame" terminators, truncated frames (zero blocks), empty and unchanged frames,
    picture sizes that leave pixels outside the 4x4 block grid."""
    rng = np.random.default_rng(5)
    out = []
    for is8 in (False, True):
        for (w, h) in ((96, 64), (320, 240), (130, 99), (64, 37)):
            nb = (w // 4) * (h // 4)
            key = synth.msv1_frame(is8, w, h, 11 + w)
            fr = [key]
            fr += [synth.msv1_frame(is8, w, h, 20 + i, skip_permille=[20, 300, 950][i % 3], mean_skip=[3, 40, 400][i % 3],
                                    mix=(30, 40, 30)) for i in range(6)]
            fr += [b"", bytes([nb & 0xFF, 0x84 + (nb >> 8)]), (b"\x00\x00" if is8 else b"\x00\x84") + b"\x1f\x80" * 9,
                   key[: len(key) // 3], key[: len(key) // 2 + 1], rng.integers(0, 256, 700, dtype=np.uint8).tobytes(), key]
            keys = [1] + [0] * (len(fr) - 2) + [1]
            out.append((is8, w, h, fr, keys, synth.random_palette(3) if is8 else None))
    return out


@pytest.mark.parametrize("flip", [False, True])
@pytest.mark.parametrize("insign", [0, 36])
def test_fused_display_store_msvideo1(flip, insign):
    """JSP_BATCH_DISPLAY: the MSVideo1 kernel stores canvas words (Manager.hx:363-381), flipped or not (Main.hx:946); pictures
    and every result flag must equal the oracle's decode followed by its display conversion, ScreenPressor streams of the same
    batch go through the separate pass, and the plain download of the batch delivers the same MSVideo1 pictures."""
    cases = _msv1_cases()
    specs = [StreamSpec(CodecType.codec_msvc8 if is8 else CodecType.codec_msvc16, w, h, 8 if is8 else 16, frames=fr, keys=k,
                        palette=pal) for is8, w, h, fr, k, pal in cases]
    spf, spk, _ = synth.sp_stream(96, 64, 5, seed=3, version=4, gop=3)
    specs.append(StreamSpec(CodecType.codec_screenpressor, 96, 64, 24, frames=spf, keys=spk))
    bd = BatchDecoder(insignificant_lines=insign, significance=True, display=True, display_flip=flip)
    bd.configure(specs)
    bd.upload(); bd.run()
    outs, flags = bd.download_display(flip=flip)
    plain, flags2 = bd.download()
    with pytest.raises(RuntimeError):
        bd.download_display(flip=not flip)                          # the flip was fixed when the batch was created
    bd.close()
    ref = BatchDecoder(insignificant_lines=insign, significance=True)
    ref.configure(specs)
    _, rflags = ref.decode_host()
    ref.close()
    assert (np.asarray(flags) == np.asarray(rflags)).all() and (np.asarray(flags2) == np.asarray(rflags)).all()
    i = 0
    for is8, w, h, fr, k, pal in cases:
        exp, ch, sg, st = O.decode_stream(O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16, w, h, 8 if is8 else 16, fr, keys=k,
                                          palette=pal, insignificant_lines=insign)
        for f in range(len(fr)):
            want = O.display_convert(exp[f], flip=flip)
            assert (outs[i] == want).all(), "%s %dx%d frame %d" % ("8-bit" if is8 else "RGB555", w, h, f)
            assert (plain[i] == want).all()
            assert bool(flags[i] & _lib.JSP_FRAME_CHANGED) == bool(ch[f])
            if not k[f]:
                assert bool(flags[i] & _lib.JSP_FRAME_SIGNIFICANT) == bool(sg[f]), (w, h, f)
            i += 1
    exp = O.decode_stream(O.CODEC_SCREENPRESSOR, 96, 64, 24, spf, keys=spk)[0]
    for f in range(len(spf)):
        assert (outs[i] == O.display_convert(exp[f], flip=flip)).all()
        assert (plain[i] == exp[f]).all()                           # ScreenPressor stays 0x00RRGGBB in HBM
        i += 1


def test_fused_display_store_end_to_end_1080p():
    """The end-to-end path of a display batch (host buffers in, canvas words out) at C2's picture size."""
    w, h = 1920, 1080
    fr = [synth.msv1_frame(False, w, h, 5)] + [synth.msv1_frame(False, w, h, 6 + i, skip_permille=850) for i in range(3)]
    spec = StreamSpec(CodecType.codec_msvc16, w, h, 16, frames=fr, keys=[1, 0, 0, 0])
    bd = BatchDecoder(display=True, display_flip=True)
    bd.configure([spec] * 3)
    outs, flags = bd.decode_host()
    bd.close()
    exp = O.decode_stream(O.CODEC_MSVC16, w, h, 16, fr, keys=[1, 0, 0, 0])[0]
    for s in range(3):
        for f in range(4):
            assert (outs[s * 4 + f] == O.display_convert(exp[f], flip=True)).all()
