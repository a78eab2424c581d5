"""BASELINE.json's configurations as parity cases (the bench line is quoted on configs[1]; the others are tested here at
their real picture sizes, with fewer streams where the config is about stream count)."""
import numpy as np
import pytest

from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def test_c1_msvideo1_pal8_320x240_300_frames():
    """configs[0]: MSVideo1 8-bit palettised 320x240, 300 frames (recipe of SURVEY.md 8d C1: key frame, then P frames
    with 85 % of the blocks skipped in runs of mean length 40, 40/40/20 % 1-/2-/8-colour blocks)."""
    w, h, n = 320, 240, 300
    pal = synth.random_palette(0xC0DEC1)
    frames = [synth.msv1_frame(True, w, h, 0xC0DEC1, mix=(40, 40, 20))]
    frames += [synth.msv1_frame(True, w, h, 0xC0DEC1 + i, skip_permille=850, mean_skip=40, mix=(40, 40, 20)) for i in range(1, n)]
    exp, ch, sg, st = O.decode_stream(O.CODEC_MSVC8, w, h, 8, frames, palette=pal, insignificant_lines=36)
    bd = BatchDecoder(insignificant_lines=36, significance=True)
    bd.configure([StreamSpec(CodecType.codec_msvc8, w, h, 8, frames=frames, palette=pal)])
    outs, flags = bd.decode_host()
    bd.close()
    for i in range(n):
        assert (outs[i] == exp[i]).all(), "frame %d" % i
        assert bool(flags[i] & _lib.JSP_FRAME_CHANGED) == bool(ch[i])
        if i:                                   # DecompressI has no significance result (IVideoCodec.hx:25)
            assert bool(flags[i] & _lib.JSP_FRAME_SIGNIFICANT) == bool(sg[i])


@pytest.mark.parametrize("version", [2, 3, 4])
def test_c4_screenpressor_1080p_inter_frames(version):
    """configs[3] at its picture size: 1920x1080 inter-frame streams (skip/copy-heavy screen content), 1 I + 7 P."""
    w, h = 1920, 1080
    specs, exp = [], []
    for s in range(2):
        frames, keys, pics = synth.sp_stream(w, h, 8, seed=0xC0DEC4 + s, version=version, change_permille=20)
        specs.append(StreamSpec(CodecType.codec_screenpressor, w, h, 24, frames=frames, keys=keys))
        exp += pics
    bd = BatchDecoder(insignificant_lines=36)
    bd.configure(specs)
    outs, flags = bd.decode_host()
    bd.close()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    for i, e in enumerate(exp):
        assert (outs[i] == e).all(), "frame %d" % i


def test_c3_screenpressor_720p_keyframes_mixed_versions():
    """configs[2] at its picture size: RGB24 1280x720 keyframe-only streams, range-coder and rANS streams mixed."""
    w, h = 1280, 720
    specs, exp = [], []
    for s in range(6):
        frames, keys, pics = synth.sp_stream(w, h, 4, seed=0xC0DEC3 + s, version=2 + s % 3, gop=1, change_permille=40)
        specs.append(StreamSpec(CodecType.codec_screenpressor, w, h, 24, frames=frames, keys=keys))
        exp += pics
    bd = BatchDecoder()
    bd.configure(specs)
    outs, flags = bd.decode_host()
    bd.close()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    for i, e in enumerate(exp):
        assert (outs[i] == e).all(), "frame %d" % i
