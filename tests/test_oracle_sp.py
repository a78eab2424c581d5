"""ScreenPressor oracle (oracle/screenpressor_oracle.c + rangecoder_oracle.c): model-level known-answer vectors and
encoder -> oracle round trips.  No independent ScreenPressor decoder exists here: parity of this
restatement with the reference is unpinned beyond these checks."""
import ctypes as C

import numpy as np
import pytest

from jsplayer_b200 import synth
from oracle import pyoracle as O


def test_kat_g5_fresh_colour_row():
    """RangeCoder.hx:84-128 on a row initialised by EntroCoders.hx:85-91: total 256, all counts 1."""
    lib = O.load()
    lib.ora_kat_rc_fresh_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    for v in (0, 1, 15, 16, 17, 200, 255):
        code = v * 0xFFFFFF + 12345                      # range/256 = 0xFFFFFF; value = code // 0xFFFFFF = v
        src = bytes([0x12, 0x00]) + code.to_bytes(4, "big") + bytes(8)
        a = np.frombuffer(src, dtype=np.uint8).copy()
        out = np.zeros(3, dtype=np.uint32)
        c = lib.ora_kat_rc_fresh_row(a.ctypes.data, a.size, 77, out.ctypes.data)
        assert c == v
        assert list(out) == [401, 416, 656]


@pytest.mark.parametrize("size", [(64, 48), (33, 17), (16, 16), (320, 240), (250, 130)])
def test_roundtrip_v2(size):
    w, h = size
    frames, keys, pics = synth.sp_stream(w, h, 8, seed=w * 31 + h, version=2, change_permille=40)
    out, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, frames, keys=keys)
    assert (st == 0).all()
    for i in range(len(frames)):
        assert (out[i] == pics[i]).all(), "frame %d" % i
    assert ch[0] == 1


def test_roundtrip_v2_gop_and_16bpp():
    w, h = 160, 96
    frames, keys, pics = synth.sp_stream(w, h, 12, seed=3, version=2, gop=4, change_permille=60)
    out, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, frames, keys=keys)
    assert all((out[i] == pics[i]).all() for i in range(12))
    frames, keys, pics = synth.sp_stream(w, h, 6, seed=4, version=2, bpp=16)
    out, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 16, frames, keys=keys)
    assert all((out[i] == pics[i]).all() for i in range(6))


def test_flat_unchanged_and_errors():
    w, h = 64, 32
    enc = synth.SPEncoder(w, h, 24, 2)
    p0 = synth.screen(w, h, 1)
    p1, mv = synth.screen_next(p0, 2, 100)
    f_i = enc.iframe(p0)
    f_p = enc.pframe(p1, p0, mv)
    f_flat = enc.flat(0x123456)
    flat_pic = np.full((h, w), 0x123456, dtype=np.int32)
    p2, mv2 = synth.screen_next(flat_pic, 3, 100)
    f_p2 = enc.pframe(p2, flat_pic, mv2)
    frames = [f_p, f_flat, f_i, f_p, b"", b"\0", f_flat, f_p2, b"\x13abc", b""]
    keys = [0, 1, 1, 0, 0, 0, 1, 0, 1, 1]
    out, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, frames, keys=keys)
    # P before any I: nothing; flat before any coded I: error (ec == null); then the coded stream
    assert list(ch[:2]) == [0, 0] and st[1] == O.ERROR_OCCURED
    assert (out[2] == p0).all() and (out[3] == p1).all()
    assert list(ch[4:6]) == [0, 0] and (out[5] == p1).all()
    assert (out[6] == flat_pic).all() and ch[6] == 1
    assert (out[7] == p2).all()
    assert st[8] == O.ERROR_OCCURED and st[9] == O.ERROR_OCCURED          # unknown head nibble / empty key frame
    assert (out[9] == p2).all()


def test_truncated_stream_reports_error():
    w, h = 96, 64
    frames, keys, pics = synth.sp_stream(w, h, 2, seed=9, version=2)
    cut = frames[0][: len(frames[0]) // 2]
    out, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, [cut], keys=[1])
    assert st[0] == O.ERROR_OCCURED
