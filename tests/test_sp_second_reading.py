"""ScreenPressor parity pin: oracle == second reading == GPU.

oracle/sp_naive.py is a second, independent, line-by-line Python reading of the reference's ANS.hx, RangeCoder.hx,
EntroCoders.hx and ScreenPressor.hx (object per context, statics and all), written without sight of oracle/*.c,
synth/ans_models.c or the CUDA kernels.  tests/golden/sp_naive_*.npz hold what it decoded from the synthetic corpus:
pictures plus the per-symbol (call kind, symbol, freq, cumFreq, total) trace (generator:
tests/golden/make_sp_naive_golden.py, which also proves the corpus covers every context-kind transition and rescale path).

  * CPU:  the C oracle reproduces every golden picture AND every golden symbol interval;
          live runs of both readings on fresh seeded streams agree (incl. the 131072-symbol rANS state reload);
  * GPU:  the CUDA path through the C ABI reproduces every golden picture.
"""
import glob
import json
import os

import numpy as np
import pytest

import synth
from oracle import pyoracle as O
from oracle import sp_naive as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "sp_naive_*.npz")))


def load(path):
    z = np.load(path)
    ln = z["frame_len"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(ln)]).astype(np.int64)
    frames = [z["data"][int(off[i]):int(off[i + 1])].tobytes() for i in range(len(ln))]
    return z, frames, [int(k) for k in z["keys"]]


def test_golden_corpus_is_present_and_covers_the_model():
    assert len(GOLDEN) >= 12
    ev = {}
    for p in GOLDEN:
        for k, v in json.loads(str(np.load(p)["events"])).items():
            ev[k] = ev.get(k, 0) + v
    for k in ("kind_1_to_4", "kind_1_to_5", "kind_1_to_2", "kind_4_to_5", "kind_5_to_6", "kind_2_to_6", "kind_2_to_3",
              "kind_3_to_7", "kind_6_to_7", "cx6_grow", "cx6_rescaleDec", "small_rescale_S4", "small_rescale_S16",
              "fixed_rebuild_256", "fixed_rebuild_6", "rc_rescale_uni", "rc_rescale_256", "rc_rescale_6"):
        assert ev.get(k, 0) > 0, k


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[9:-4] for p in GOLDEN])
def test_oracle_matches_second_reading_golden(path):
    z, frames, keys = load(path)
    w, h, bpp = int(z["width"]), int(z["height"]), int(z["bpp"])
    out, ch, sg, st, tr = O.decode_stream_traced(O.CODEC_SCREENPRESSOR, w, h, bpp, frames, keys=keys)
    assert (st == 0).all()
    assert (out == z["pictures"]).all()
    assert list(ch) == list(z["changed"])
    assert list(sg) == list(z["significant"])
    g = z["trace"]
    assert tr.shape == g.shape, "symbol count %d vs %d" % (tr.shape[0], g.shape[0])
    bad = np.nonzero((tr != g).any(axis=1))[0]
    assert bad.size == 0, "first differing symbol %d: oracle %s, second reading %s" % (bad[0], tr[bad[0]], g[bad[0]])


def _live(w, h, bpp, frames, keys):
    pics, ch, sg, traces, info = N.decode_stream(w, h, bpp, frames, trace=True)
    out, och, osg, ost, otr = O.decode_stream_traced(O.CODEC_SCREENPRESSOR, w, h, bpp, frames, keys=keys)
    assert (ost == 0).all()
    for f in range(len(frames)):
        if pics[f] is not None:
            assert (np.array(pics[f], dtype=np.int64).astype(np.int32).reshape(h, w) == out[f]).all(), "frame %d" % f
    assert list(ch) == [bool(v) for v in och] and list(sg) == [bool(v) for v in osg]
    ntr = np.array([t for fr in traces for t in fr], dtype=np.int32).reshape(-1, 5)
    assert ntr.shape == otr.shape and (ntr == otr).all()
    return info


@pytest.mark.parametrize("version", [2, 3, 4])
def test_live_second_reading_on_fresh_streams(version):
    """Seeds and sizes that are NOT in the committed corpus."""
    rng = np.random.Generator(np.random.PCG64(1000 + version))
    for _ in range(6):
        w, h = int(rng.integers(5, 140)), int(rng.integers(5, 90))
        bpp = 16 if rng.integers(0, 4) == 0 and version != 4 else 24
        frames, keys, _ = synth.sp_stream(w, h, int(rng.integers(2, 7)), seed=int(rng.integers(1, 1 << 30)), version=version,
                                          gop=int(rng.integers(0, 4)), change_permille=int(rng.integers(10, 300)), bpp=bpp)
        _live(w, h, bpp, frames, keys)


def test_live_second_reading_rans_state_reload():
    """EntroCoders.hx:249-253: both readings re-read the rANS state after Rans.B = 131072 symbols."""
    w, h = 400, 260
    enc = synth.SPEncoder(w, h, 24, 4)
    px = synth.noise(w, h, 11, ncolors=(3, 9, 30))
    N.EVENTS.clear()
    _live(w, h, 24, [enc.iframe(px)], [1])
    assert N.EVENTS["rans_reinit"] >= 1


def test_second_reading_known_answers():
    """SURVEY.md Appendix G5/G6 against the second reading (the oracle's own KATs are in test_oracle_sp.py)."""
    ec = N.EntroCoderRC()
    ec.preinit()
    ec.renewI()
    v = 200
    code = v * 0xFFFFFF + 12345
    ec.decodeBegin(N.U8(list(bytes([0x12, 0x00]) + code.to_bytes(4, "big") + bytes(8))), 1)
    assert ec.decodeClr(77) == v
    row = ec.cntab.rows[77]
    assert (row[17 + v], row[v >> 4], row[16]) == (401, 416, 656)
    cx = N.Context()
    assert cx.decode(0) is False
    cx.update(77)
    assert cx.u[0] == 1
    cx.update(77)
    assert cx.u[0] == 4 and cx.u[1].d == 1 and cx.u[1].freqs[0] == 100
    t = N.FixedSizeRansCtx(256)
    t.renew()
    assert (t.freqs[6], t.freqs[7], t.cnts[3], t.cntsum, t.decTable[5]) == (16, 48, 8, 2048, 40)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[9:-4] for p in GOLDEN])
def test_gpu_matches_second_reading_golden(path):
    from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
    z, frames, keys = load(path)
    w, h, bpp = int(z["width"]), int(z["height"]), int(z["bpp"])
    bd = BatchDecoder(significance=True)
    bd.configure([StreamSpec(CodecType.codec_screenpressor, w, h, bpp, frames=frames, keys=keys)])
    outs, flags = bd.decode_host()
    bd.close()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    for f in range(len(frames)):
        if z["have_picture"][f] or f > 0:
            assert (outs[f] == z["pictures"][f]).all(), "frame %d" % f
        assert bool(flags[f] & _lib.JSP_FRAME_CHANGED) == bool(z["changed"][f]), "changed flag of frame %d" % f
        assert bool(flags[f] & _lib.JSP_FRAME_SIGNIFICANT) == bool(z["significant"][f]), "significant flag of frame %d" % f


# ---------------------------------------------------------------- hand-crafted syntax (tests/sp_crafted.py) ----
from sp_crafted import CRAFTED                                   # noqa: E402


@pytest.mark.parametrize("version", [2, 3, 4])
@pytest.mark.parametrize("name", sorted(CRAFTED))
def test_crafted_streams_oracle_matches_second_reading(name, version):
    """Streams no sane encoder emits (out-of-picture predictors, zero-length runs, predictor 3 in an I frame, wild motion
    vectors, sub-rectangles past the right edge), encoded by the second reading's own symbol-level encoder."""
    w, h, bpp, frames, keys = CRAFTED[name](version)
    _live(w, h, bpp, frames, keys)


@pytest.mark.gpu
@pytest.mark.parametrize("version", [2, 3, 4])
@pytest.mark.parametrize("name", sorted(CRAFTED))
def test_crafted_streams_gpu_matches_second_reading(name, version):
    from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
    w, h, bpp, frames, keys = CRAFTED[name](version)
    pics, ch, sg, _, _ = N.decode_stream(w, h, bpp, frames)
    bd = BatchDecoder(significance=True)
    bd.configure([StreamSpec(CodecType.codec_screenpressor, w, h, bpp, frames=frames, keys=keys)])
    outs, flags = bd.decode_host()
    bd.close()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    for f in range(len(frames)):
        assert (outs[f] == np.array(pics[f], dtype=np.int64).astype(np.int32).reshape(h, w)).all(), "frame %d" % f


def _truncated_flat_stream(bpp):
    w, h = 32, 16
    frames, keys, _ = synth.sp_stream(w, h, 1, seed=3, version=2, bpp=bpp)
    for tail in (b"", b"\x5a", b"\x5a\xa5", b"\x5a\xa5\x3c", b"\x5a\xa5\x3c\x77"):
        frames.append(b"\x11" + tail)
        keys.append(1)
    return w, h, frames, keys


@pytest.mark.parametrize("bpp", [16, 24])
def test_truncated_flat_frames_oracle_matches_second_reading(bpp):
    """A flat key frame shorter than its colour: `src[1] * 256` with src[1] undefined is NaN and the fill colour 0
    (ScreenPressor.hx:136-147) -- not "the missing byte counts as 0"."""
    w, h, frames, keys = _truncated_flat_stream(bpp)
    pics, ch, sg, _, _ = N.decode_stream(w, h, bpp, frames)
    out, och, osg, ost = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, bpp, frames, keys=keys)
    assert (ost == 0).all()
    for f in range(len(frames)):
        assert (np.array(pics[f], dtype=np.int64).astype(np.int32).reshape(h, w) == out[f]).all(), "frame %d" % f


@pytest.mark.gpu
@pytest.mark.parametrize("bpp", [16, 24])
def test_truncated_flat_frames_gpu_matches_second_reading(bpp):
    from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType
    w, h, frames, keys = _truncated_flat_stream(bpp)
    pics, _, _, _, _ = N.decode_stream(w, h, bpp, frames)
    bd = BatchDecoder()
    bd.configure([StreamSpec(CodecType.codec_screenpressor, w, h, bpp, frames=frames, keys=keys)])
    outs, flags = bd.decode_host()
    bd.close()
    for f in range(len(frames)):
        assert (outs[f] == np.array(pics[f], dtype=np.int64).astype(np.int32).reshape(h, w)).all(), "frame %d" % f
