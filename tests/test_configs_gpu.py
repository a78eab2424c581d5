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


def test_c5_mixed_4k_corpus_full_gops(tmp_path):
    """configs[4] at its picture size: 3840x2160 AVI files (one MSVideo1 RGB555, one ScreenPressor v4, one v2), a full GOP of
    16 frames each plus the next key frame, through the AVI indexer and the GOP splitter, bit-exact against the oracle
    decoding every file as one stream."""
    from jsplayer_b200 import avi
    from synth.avi import write_avi
    w, h, gop, n = 3840, 2160, 16, 17
    files = []
    frames = [synth.msv1_frame(False, w, h, 0xC0DEC5 * 64 + f, skip_permille=0 if f % gop == 0 else 850) for f in range(n)]
    keys = [1 if f % gop == 0 else 0 for f in range(n)]
    files.append(("m.avi", O.CODEC_MSVC16, 16, b"CRAM", frames, keys))
    for i, version in enumerate((4, 2)):
        frames, keys, _ = synth.sp_stream(w, h, n, seed=0xC0DEC5 + 1 + i, version=version, gop=gop, change_permille=20)
        files.append(("s%d.avi" % version, O.CODEC_SCREENPRESSOR, 24, b"SCPR", frames, keys))
    streams = []
    for name, codec, bpp, four, frames, keys in files:
        path = str(tmp_path / name)
        write_avi(path, w, h, bpp, four, frames, keys)
        streams.append(avi.load_avi(path, pinned=True))
    specs, where = avi.gop_specs(streams)
    assert len(specs) == 2 * len(files)
    bd = BatchDecoder(insignificant_lines=36)
    bd.configure(specs)
    outs, flags = bd.decode_host()
    bd.close()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    exp = [O.decode_stream(codec, w, h, bpp, frames, keys=keys, insignificant_lines=36)[0] for _, codec, bpp, _, frames, keys in files]
    i = 0
    for (fi, lo, hi) in where:
        for f in range(lo, hi):
            assert (outs[i] == exp[fi][f]).all(), "file %d frame %d" % (fi, f)
            i += 1
    assert i == n * len(files)


def test_c4_at_512_streams_spot_checked():
    """configs[3] at its stream count: 512 ScreenPressor 1080p streams of 1 I + 31 P frames in ONE batch (136 GB of pictures
    stay in HBM; 16 distinct streams, v2 and v4 alternating, repeated -- every copy decodes with its own model state).
    Spot check: six streams spread over the batch, all 32 frames, against the oracle; no frame of the batch may fail."""
    w, h, n_streams, n_frames, distinct = 1920, 1080, 512, 32, 16
    base = [synth.sp_stream(w, h, n_frames, seed=0xC0DEC4 + i, version=(2, 4)[i % 2], change_permille=20)[:2] for i in range(distinct)]
    specs = [StreamSpec(CodecType.codec_screenpressor, w, h, 24, frames=base[i % distinct][0], keys=base[i % distinct][1])
             for i in range(n_streams)]
    bd = BatchDecoder(insignificant_lines=36)
    bd.configure(specs, pinned=True)
    bd.upload(); bd.run(); bd.sync()
    check = [0, 1, 130, 259, 388, 511]
    outs = [None] * bd.n_frames
    for s in check:
        for f in range(n_frames):
            outs[s * n_frames + f] = np.empty((h, w), dtype=np.int32)
    _, flags = bd.download(outs)
    bd.close()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    for s in check:
        fr, keys = base[s % distinct]
        exp = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, fr, keys=keys, insignificant_lines=36)[0]
        for f in range(n_frames):
            assert (outs[s * n_frames + f] == exp[f]).all(), "stream %d frame %d" % (s, f)
