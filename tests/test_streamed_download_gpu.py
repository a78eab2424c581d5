"""GPU parity for the level-streamed end-to-end path: with page-locked destination pictures, jsp_batch_decode_host
copies out the pictures each launch finishes while later dependency levels still decode.  Results must be identical
to the back-to-back path (pageable destinations) and to the CPU oracle."""
import numpy as np
import pytest

from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu
SP = CodecType.codec_screenpressor


def decode_pinned(specs, insign=0):
    bd = BatchDecoder(insignificant_lines=insign, significance=True)
    bd.configure(specs, pinned=True)
    outs = bd.alloc_outputs(pinned=True)
    for o in outs:
        o[...] = -1                      # anything left unwritten would show
    outs, flags = bd.decode_host(outs)
    outs = [o.copy() for o in outs]
    bd.close()
    return outs, flags


def sp_mixed_stream(w, h, version, seed):
    enc = synth.SPEncoder(w, h, 24, version)
    p0 = synth.screen(w, h, seed)
    p1, mv = synth.screen_next(p0, seed + 1, 100)
    f_i = enc.iframe(p0)
    f_p = enc.pframe(p1, p0, mv)
    f_flat = enc.flat(0x123456)
    flat_pic = np.full((h, w), 0x123456, dtype=np.int32)
    p2, mv2 = synth.screen_next(flat_pic, seed + 2, 100)
    f_p2 = enc.pframe(p2, flat_pic, mv2)
    frames = [f_p, f_flat, f_i, f_p, b"", b"\0", f_flat, f_flat, f_p2, b"\x13abc", b""]
    keys = [0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 1]
    return frames, keys


def test_screenpressor_chains_stream_out_level_by_level():
    specs, exp = [], []
    for s, (w, h, version) in enumerate([(64, 32, 2), (80, 48, 4), (33, 17, 3), (160, 96, 2), (160, 96, 4)]):
        if s < 3:
            frames, keys = sp_mixed_stream(w, h, version, 10 + s)
        else:
            frames, keys, _ = synth.sp_stream(w, h, 12, seed=s, version=version, gop=5, change_permille=60)
        specs.append(StreamSpec(SP, w, h, 24, frames=frames, keys=keys))
        exp.append(O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, frames, keys=keys, insignificant_lines=16))
    outs, flags = decode_pinned(specs, insign=16)
    i = 0
    for s, (pics, ch, sg, st) in enumerate(exp):
        for f in range(len(pics)):
            assert bool(flags[i] & _lib.JSP_FRAME_ERROR) == (st[f] != 0), (s, f)
            assert (outs[i] == pics[f]).all(), (s, f)
            if st[f] == 0:
                assert bool(flags[i] & _lib.JSP_FRAME_CHANGED) == bool(ch[f]), (s, f)
            i += 1
    assert i == len(outs)


def test_msvideo1_p_chains_and_mixed_codecs_stream_out():
    specs, exp = [], []
    for s in range(10):
        is8 = bool(s & 1)
        w, h = [(64, 48), (320, 240), (100, 60)][s % 3]
        pal = synth.random_palette(s) if is8 else None
        frames = [synth.msv1_frame(is8, w, h, 100 + s)] + [
            synth.msv1_frame(is8, w, h, 200 + s * 10 + i, skip_permille=150, mean_skip=12) for i in range(1 + s % 4)] + [b""]
        keys = [1] + [0] * (len(frames) - 1)
        codec = O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16
        specs.append(StreamSpec(CodecType.codec_msvc8 if is8 else CodecType.codec_msvc16, w, h, 8 if is8 else 16,
                                frames=frames, keys=keys, palette=pal))
        exp.append(O.decode_stream(codec, w, h, 8 if is8 else 16, frames, keys=keys, palette=pal)[0])
    frames, keys, _ = synth.sp_stream(96, 64, 6, seed=77, version=2, change_permille=50)
    specs.append(StreamSpec(SP, 96, 64, 24, frames=frames, keys=keys))
    exp.append(O.decode_stream(O.CODEC_SCREENPRESSOR, 96, 64, 24, frames, keys=keys)[0])
    outs, flags = decode_pinned(specs)
    i = 0
    for s, pics in enumerate(exp):
        for f in range(len(pics)):
            e = np.asarray(pics[f])
            bh, bw = (e.shape[0] & ~3, e.shape[1] & ~3) if s < 10 else e.shape
            assert (outs[i][:bh, :bw] == e[:bh, :bw]).all(), (s, f)
            i += 1
    assert i == len(outs)


def test_demoted_key_frame_is_downloaded_again():
    w, h = 64, 48
    frames = [synth.msv1_frame(False, w, h, 1), synth.msv1_frame(False, w, h, 2, skip_permille=300),
              synth.msv1_frame(False, w, h, 3, skip_permille=300)]
    exp_a = O.decode_stream(O.CODEC_MSVC16, w, h, 16, frames, keys=[1, 1, 1])[0]
    exp_b = O.decode_stream(O.CODEC_MSVC16, w, h, 16, frames, keys=[1, 0, 0])[0]
    # two streams so that the plan has several launches; every frame of the first claims to be a key frame
    outs, flags = decode_pinned([StreamSpec(CodecType.codec_msvc16, w, h, 16, frames=frames, keys=[1, 1, 1]),
                                 StreamSpec(CodecType.codec_msvc16, w, h, 16, frames=frames, keys=[1, 0, 0])])
    for i in range(6):
        assert (outs[i] == (exp_a if i < 3 else exp_b)[i % 3]).all(), i
        assert not (flags[i] & _lib.JSP_FRAME_ERROR)


@pytest.mark.parametrize("version", [2, 4])
def test_single_launch_pictures_leave_while_the_launch_runs(version):
    """Key frames only: one launch, whose pictures the host copies out frame by frame as the kernel flags them complete
    (short frames long before the longest one ends).  Repeated, to give an ordering bug a chance to show."""
    specs, exp = [], []
    for s, (w, h) in enumerate([(320, 240), (64, 48), (640, 360), (33, 17), (1280, 720), (100, 60)]):
        frames, keys, pics = synth.sp_stream(w, h, 3, seed=50 + s, version=version, gop=1, change_permille=60)
        specs.append(StreamSpec(SP, w, h, 24, frames=frames, keys=keys))
        exp += pics
    for rep in range(3):
        outs, flags = decode_pinned(specs)
        for i, e in enumerate(exp):
            assert (outs[i] == e).all(), (rep, i)
            assert not (flags[i] & _lib.JSP_FRAME_ERROR)
