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


def test_c4_at_512_streams_every_frame():
    """configs[3] at its stream count: 512 DISTINCT ScreenPressor 1080p streams of 1 I + 31 P frames (v2 and v4 alternating) in
    ONE batch -- 136 GB of pictures stay in HBM -- and every one of the 16 384 pictures compared with the oracle (downloaded
    16 streams at a time; the oracle decodes on 16 host threads)."""
    from concurrent.futures import ThreadPoolExecutor
    w, h, n_streams, n_frames = 1920, 1080, 512, 32

    def make(i):
        return synth.sp_stream(w, h, n_frames, seed=0xC4000 + i, version=(2, 4)[i % 2], change_permille=20)[:2]
    with ThreadPoolExecutor(max_workers=16) as ex:
        streams = list(ex.map(make, range(n_streams)))
    bd = BatchDecoder(insignificant_lines=36)
    bd.configure([StreamSpec(CodecType.codec_screenpressor, w, h, 24, frames=fr, keys=k) for fr, k in streams], pinned=True)
    bd.upload(); bd.run(); bd.sync()
    flags = bd.results()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    group = 16
    bufs = [np.empty((h, w), dtype=np.int32) for _ in range(group * n_frames)]
    for lo in range(0, n_streams, group):
        outs = [None] * bd.n_frames
        for i in range(group * n_frames):
            outs[lo * n_frames + i] = bufs[i]
        bd.download(outs)

        def check(s):
            fr, keys = streams[lo + s]
            exp = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, fr, keys=keys, insignificant_lines=36)[0]
            return [f for f in range(n_frames) if not (bufs[s * n_frames + f] == exp[f]).all()]
        with ThreadPoolExecutor(max_workers=16) as ex:
            bad = list(ex.map(check, range(group)))
        for s, b in enumerate(bad):
            assert not b, "stream %d frames %s" % (lo + s, b)
    bd.close()


def _oracle_all(specs_args, threads=16):
    """Oracle pictures of many streams, one stream per worker thread (the C oracle releases the GIL inside ctypes calls)."""
    from concurrent.futures import ThreadPoolExecutor

    def one(a):
        codec, w, h, bpp, frames, keys = a
        return O.decode_stream(codec, w, h, bpp, frames, keys=keys, insignificant_lines=36)[0]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(one, specs_args))


def test_c2_full_batch_every_frame_against_the_oracle():
    """configs[1] at its full size -- 1024 MSVideo1 RGB555 1920x1080 key frames in ONE batch (the bench's headline launch) -- and
    EVERY picture compared with the oracle, not a sample: pictures are downloaded 64 at a time."""
    from concurrent.futures import ThreadPoolExecutor
    w, h, n = 1920, 1080, 1024
    with ThreadPoolExecutor(max_workers=16) as ex:
        frames = list(ex.map(lambda i: synth.msv1_frame(False, w, h, 0xC0DEC2 + i, mix=(25, 50, 25)), range(n)))
    bd = BatchDecoder(insignificant_lines=36)
    bd.configure([StreamSpec(CodecType.codec_msvc16, w, h, 16, frames=[f]) for f in frames], pinned=True)
    bd.upload(); bd.run(); bd.sync()
    flags = bd.results()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    assert (flags & _lib.JSP_FRAME_CHANGED).all()
    bufs = [np.empty((h, w), dtype=np.int32) for _ in range(64)]
    for lo in range(0, n, 64):
        outs = [None] * n
        for i in range(64):
            outs[lo + i] = bufs[i]
        bd.download(outs)
        exp = _oracle_all([(O.CODEC_MSVC16, w, h, 16, [frames[lo + i]], None) for i in range(64)])
        for i in range(64):
            assert (bufs[i] == exp[i][0]).all(), "frame %d" % (lo + i)
    bd.close()


def test_c3_full_batch_every_frame_against_the_oracle():
    """configs[2] at its full size -- 256 ScreenPressor 1280x720 streams of 4 I frames, range-coder and rANS streams alternating,
    256 DISTINCT streams (the bench repeats 32) in one batch of 1024 independent frames -- every picture against the oracle."""
    from concurrent.futures import ThreadPoolExecutor
    w, h, n_streams = 1280, 720, 256

    def make(i):
        fr, keys, _ = synth.sp_stream(w, h, 4, seed=0xC3000 + i, version=(2, 4, 3, 2)[i % 4], gop=1, change_permille=40)
        return fr, keys
    with ThreadPoolExecutor(max_workers=16) as ex:
        streams = list(ex.map(make, range(n_streams)))
    bd = BatchDecoder(insignificant_lines=36)
    bd.configure([StreamSpec(CodecType.codec_screenpressor, w, h, 24, frames=fr, keys=k) for fr, k in streams], pinned=True)
    bd.upload(); bd.run(); bd.sync()
    flags = bd.results()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    bufs = [np.empty((h, w), dtype=np.int32) for _ in range(128)]
    for lo in range(0, n_streams, 32):                       # 32 streams = 128 pictures at a time
        outs = [None] * (4 * n_streams)
        for i in range(128):
            outs[4 * lo + i] = bufs[i]
        bd.download(outs)
        exp = _oracle_all([(O.CODEC_SCREENPRESSOR, w, h, 24, streams[lo + s][0], streams[lo + s][1]) for s in range(32)])
        for s in range(32):
            for f in range(4):
                assert (bufs[4 * s + f] == exp[s][f]).all(), "stream %d frame %d" % (lo + s, f)
    bd.close()
