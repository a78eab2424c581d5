"""ScreenPressor oracle (oracle/screenpressor_oracle.c + rangecoder_oracle.c): model-level known-answer vectors and
encoder -> oracle round trips.  No independent ScreenPressor decoder exists here: parity of this
restatement with the reference is unpinned beyond these checks."""
import ctypes as C

import numpy as np
import pytest

import synth
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


# ---------------------------------------------------------------- rANS streams (v3 / v4) ----
def test_kat_g6_fresh_ans_models():
    """ANS.hx:226-238,831-835 (second equal symbol -> Cx4, freqs [100], d 1) and
    FixedSizeRansCtx(256).renew() (:128-144): freq 16, cum 16 i, count 8, cntsum 2048, decTable[k] = 8 k."""
    lib = O.load()
    out = (C.c_int * 10)()
    lib.ora_kat_ans(77, out)
    assert list(out) == [0, 1, 4, 1, 100, 16, 48, 8, 2048, 40]


@pytest.mark.parametrize("version", [3, 4])
@pytest.mark.parametrize("size", [(64, 48), (33, 17), (320, 240)])
def test_roundtrip_ans(version, size):
    w, h = size
    frames, keys, pics = synth.sp_stream(w, h, 8, seed=w * 31 + h, version=version, change_permille=40)
    assert frames[0][0] == ((version - 1) << 4 | 2)
    out, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, frames, keys=keys)
    assert (st == 0).all()
    for i in range(len(frames)):
        assert (out[i] == pics[i]).all(), "frame %d" % i


def many_symbol_picture(w, h, seed):
    """Noise bands plus a first band whose red channel walks through 100 distinct values before repeating while
    green = blue = 0: one colour context meets > 64 distinct symbols before its first repeat (Cx3 -> Cx7)."""
    px = synth.noise(w, h, seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    perm = rng.permutation(256)[:100]
    seq = np.concatenate([perm, perm, perm])
    n = min(w * 4, seq.size)
    px.reshape(-1)[:n] = seq[:n]
    return px


@pytest.mark.parametrize("version", [3, 4])
def test_roundtrip_ans_all_context_kinds(version):
    """Content that drives the colour contexts through every kind transition of ANS.hx:785-860."""
    w, h = 384, 256
    synth.ans_transitions(reset=True)
    enc = synth.SPEncoder(w, h, 24, version)
    px = many_symbol_picture(w, h, 5)
    f0 = enc.iframe(px)
    nxt = synth.noise(w, h, 6)
    nxt[: h // 2] = px[: h // 2]
    f1 = enc.pframe(nxt, px)
    tr = synth.ans_transitions()
    assert all(tr[k] > 0 for k in synth.ANS_TRANSITIONS), tr
    out, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, [f0, f1], keys=[1, 0])
    assert (st == 0).all()
    assert (out[0] == px).all() and (out[1] == nxt).all()


def test_ans_state_reload_every_131072_symbols():
    """EntroCoders.hx:249-253: the decoder re-reads the rANS state every Rans.B symbols."""
    w, h = 512, 300                                        # > 131072 symbols in one noisy I frame
    enc = synth.SPEncoder(w, h, 24, 4)
    px = synth.noise(w, h, 11, ncolors=(3, 9, 30))
    f0 = enc.iframe(px)
    out, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, [f0], keys=[1])
    assert st[0] == 0 and (out[0] == px).all()
