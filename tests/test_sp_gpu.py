"""GPU parity: ScreenPressor decoded by the CUDA path through the C ABI must be bit-exact against the CPU oracle --
pictures, `changed`, `significant_changes` and error status."""
import numpy as np
import pytest

from jsplayer_b200 import BatchDecoder, StreamSpec, ScreenPressor, CodecType, DecoderState, _lib
import synth
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu
SP = CodecType.codec_screenpressor


def gpu_decode(specs, insign=0):
    bd = BatchDecoder(insignificant_lines=insign, significance=True)
    bd.configure(specs)
    outs, flags = bd.decode_host()
    bd.close()
    return outs, flags


def check(w, h, bpp, frames, keys, insign=0):
    exp, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, bpp, frames, keys=keys, insignificant_lines=insign)
    outs, flags = gpu_decode([StreamSpec(SP, w, h, bpp, frames=frames, keys=keys)], insign)
    for i in range(len(frames)):
        err = bool(flags[i] & _lib.JSP_FRAME_ERROR)
        assert err == (st[i] != 0), "error status of frame %d" % i
        # a failed frame shows the previous picture (P) or nothing (I), exactly as the oracle materialises it
        assert (outs[i] == exp[i]).all(), "frame %d differs" % i
        if not err:
            assert bool(flags[i] & _lib.JSP_FRAME_CHANGED) == bool(ch[i]), "changed flag of frame %d" % i
            assert bool(flags[i] & _lib.JSP_FRAME_SIGNIFICANT) == bool(sg[i]), "significant flag of frame %d" % i


@pytest.mark.parametrize("version", [2, 3, 4])
@pytest.mark.parametrize("size", [(64, 48), (33, 17), (16, 16), (20, 9), (320, 240), (250, 130), (1280, 720)])
def test_iframes_and_pframes(size, version):
    w, h = size
    n = 4 if w >= 1280 else 8
    frames, keys, pics = synth.sp_stream(w, h, n, seed=w * 31 + h, version=version, change_permille=40)
    check(w, h, 24, frames, keys)
    check(w, h, 24, frames, keys, insign=36)


def test_v2_gop_16bpp_and_long_stream():
    w, h = 160, 96
    frames, keys, pics = synth.sp_stream(w, h, 24, seed=3, version=2, gop=6, change_permille=60)
    check(w, h, 24, frames, keys, insign=16)
    frames, keys, pics = synth.sp_stream(w, h, 6, seed=4, version=2, bpp=16)
    check(w, h, 16, frames, keys)


def test_flat_unchanged_and_error_frames():
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
    frames = [f_p, f_flat, f_i, f_p, b"", b"\0", f_flat, f_flat, f_p2, b"\x13abc", b""]
    keys = [0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 1]
    check(w, h, 24, frames, keys)


@pytest.mark.parametrize("version", [2, 4])
def test_truncated_and_garbage_streams(version):
    w, h = 96, 64
    frames, keys, pics = synth.sp_stream(w, h, 3, seed=9, version=version, change_permille=80)
    rng = np.random.default_rng(4)
    for cut in (len(frames[0]) // 2, 7, 2):
        check(w, h, 24, [frames[0][:cut]], [1])
    check(w, h, 24, [frames[0], frames[1][: len(frames[1]) // 2], frames[2]], [1, 0, 0])
    garbage = bytes([0x02 | (version - 1) << 4]) + rng.integers(0, 256, 3000, dtype=np.uint8).tobytes()
    check(w, h, 24, [garbage], [1])
    check(w, h, 24, [frames[0], bytes([1]) + rng.integers(0, 256, 500, dtype=np.uint8).tobytes()], [1, 0])


def test_many_streams():
    specs, exp = [], []
    for s in range(20):
        w, h = [(64, 48), (320, 240), (100, 60), (640, 360)][s % 4]
        frames, keys, pics = synth.sp_stream(w, h, 2 + s % 4, seed=100 + s, version=2 + s % 3, change_permille=30)
        specs.append(StreamSpec(SP, w, h, 24, frames=frames, keys=keys))
        exp += pics
    outs, flags = gpu_decode(specs)
    for i, e in enumerate(exp):
        assert (outs[i] == e).all(), i
        assert not (flags[i] & _lib.JSP_FRAME_ERROR)


@pytest.mark.parametrize("version", [2, 3])
def test_repeated_runs_are_idempotent(version):
    w, h = 320, 240
    frames, keys, pics = synth.sp_stream(w, h, 5, seed=77, version=version)
    bd = BatchDecoder()
    bd.configure([StreamSpec(SP, w, h, 24, frames=frames, keys=keys)])
    bd.upload()
    for _ in range(3):
        bd.run()
    outs, flags = bd.download()
    bd.close()
    for i in range(5):
        assert (outs[i] == pics[i]).all()


@pytest.mark.parametrize("version", [2, 4])
def test_per_stream_dropin_matches_oracle(version):
    w, h = 200, 120
    frames, keys, pics = synth.sp_stream(w, h, 10, seed=5, version=version, gop=5, change_permille=50)
    frames.insert(3, b"\0"); keys.insert(3, 0)
    mine = ScreenPressor(w, h, 24)
    ora = O.OracleCodec(O.CODEC_SCREENPRESSOR, w, h, 24)
    mine.Preinit(36); ora.Preinit(36)
    assert not mine.NeedsIndex()
    bufs_m = [np.zeros(w * h, dtype=np.int32) for _ in range(3)]
    bufs_o = [np.zeros(w * h, dtype=np.int32) for _ in range(3)]
    for i, f in enumerate(frames):
        assert mine.IsKeyFrame(f) == ora.IsKeyFrame(f) == bool(keys[i])
        dm = next(b for b in bufs_m if b is not mine.PreviousFrame())
        do = next(b for b in bufs_o if b is not ora.PreviousFrame())
        if keys[i]:
            assert mine.DecompressI(f, dm) == DecoderState.zero_state
            assert ora.DecompressI(f, do) == 0
            pm, po, sm, so = mine.PreviousFrame(), ora.PreviousFrame(), False, False
        else:
            r = mine.DecompressP(f, dm)
            po, so = ora.DecompressP(f, do)
            pm, sm = r.data_pnt, r.significant_changes
        assert (pm is dm) == (po is do), "frame %d: data_pnt identity" % i
        assert sm == so, "frame %d: significant_changes" % i
        assert (pm == po).all(), "frame %d: picture" % i



# ---------------------------------------------------------------- rANS specifics (v3 / v4) ----
def many_symbol_picture(w, h, seed):
    """see tests/test_oracle_sp.py: drives one colour context through Cx1 -> Cx2 -> Cx3 -> Cx7"""
    px = synth.noise(w, h, seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    perm = rng.permutation(256)[:100]
    seq = np.concatenate([perm, perm, perm])
    n = min(w * 4, seq.size)
    px.reshape(-1)[:n] = seq[:n]
    return px


@pytest.mark.parametrize("version", [3, 4])
def test_ans_all_context_kinds(version):
    """Noisy content drives the colour contexts through every kind transition of ANS.hx:785-860."""
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
    check(w, h, 24, [f0, f1], [1, 0])


def test_ans_state_reload_every_131072_symbols():
    w, h = 512, 300
    enc = synth.SPEncoder(w, h, 24, 4)
    px = synth.noise(w, h, 11, ncolors=(3, 9, 30))
    check(w, h, 24, [enc.iframe(px)], [1])


def test_ans_noise_many_sizes():
    for i, (w, h) in enumerate([(64, 64), (130, 70), (257, 33)]):
        for version in (3, 4):
            enc = synth.SPEncoder(w, h, 24, version)
            a = synth.noise(w, h, 20 + i, ncolors=(2, 5, 17, 70, 256))
            b = synth.noise(w, h, 40 + i, ncolors=(256, 3))
            b[::2] = a[::2]
            check(w, h, 24, [enc.iframe(a), enc.pframe(b, a), enc.iframe(b)], [1, 0, 1])


@pytest.mark.parametrize("version", [2, 4])
def test_independent_segments_run_concurrently(version):
    """Every coded I frame resets all models (ScreenPressor.hx:163, EntroCoders.hx:81-130, 216-227): the batcher decodes
    the segments of one stream concurrently with one model-state slot each (more segments than slots here)."""
    w, h = 96, 80
    frames, keys, pics = synth.sp_stream(w, h, 40, seed=21, version=version, gop=2, change_permille=80)
    check(w, h, 24, frames, keys)
    frames, keys, pics = synth.sp_stream(w, h, 21, seed=22, version=version, gop=1)
    enc = synth.SPEncoder(w, h, 24, version)
    frames.insert(5, enc.flat(0x0A0B0C)); keys.insert(5, 1)      # a model-resetting flat frame is a segment of its own
    frames.insert(6, b"\0"); keys.insert(6, 0)
    check(w, h, 24, frames, keys, insign=16)


def test_16bpp_corrupt_stream_stays_in_bounds():
    """16 bpp v2 streams index colour contexts with whole channel bytes (SC_CXSHIFT 0, ScreenPressor.hx:59,200-202):
    corrupt symbols above 31 push cx + cx1 past a channel's 4096 contexts.  Defined behaviour: wrap + failed frame."""
    w, h = 64, 48
    frames, keys, pics = synth.sp_stream(w, h, 3, seed=8, version=2, bpp=16, change_permille=200)
    rng = np.random.default_rng(12)
    for trial in range(6):
        bad = bytearray(frames[0])
        for _ in range(6):
            bad[int(rng.integers(8, len(bad)))] ^= int(rng.integers(1, 256))
        check(w, h, 16, [bytes(bad), frames[1], frames[2]], keys)
        badp = bytearray(frames[1])
        for _ in range(3):
            badp[int(rng.integers(2, len(badp)))] ^= int(rng.integers(1, 256))
        check(w, h, 16, [frames[0], bytes(badp), frames[2]], keys)


def test_symbol_count_matches_the_oracle():
    """jsp_batch_symbols: the device-side count of entropy-coded symbols equals the oracle's (both coders)."""
    import ctypes as C
    lib = O.load()
    lib.ora_symbol_count.restype = C.c_ulonglong
    lib.ora_symbol_count.argtypes = [C.c_int]
    for version in (2, 4):
        w, h = 160, 120
        frames, keys, pics = synth.sp_stream(w, h, 6, seed=40 + version, version=version, gop=3, change_permille=80)
        lib.ora_symbol_count(1)
        O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, 24, frames, keys=keys)
        want = lib.ora_symbol_count(1)
        bd = BatchDecoder()
        bd.configure([StreamSpec(SP, w, h, 24, frames=frames, keys=keys)])
        bd.upload(); bd.run(); bd.sync()
        got = bd.symbols()
        bd.run(); bd.sync()
        assert bd.symbols() == got                                  # per run, not cumulative
        bd.close()
        assert got == want and got > 0


@pytest.mark.parametrize("version", [2, 3, 4])
def test_corrupt_p_frames_match_the_oracle(version):
    """Bit flips in P frames: rectangles that degenerate or leave the picture, runs that overflow their rectangle,
    motion vectors pointing outside -- whatever the decoder then does must equal the oracle, frame by frame (the
    shared-memory block tile of well-formed rectangles and the global-memory path of the others mix freely)."""
    rng = np.random.default_rng(100 + version)
    for (w, h) in ((100, 60), (33, 17), (64, 48), (16, 16), (250, 130)):
        frames, keys, pics = synth.sp_stream(w, h, 4, seed=w + version, version=version, change_permille=250)
        for trial in range(12):
            bad = [frames[0]]
            for f in frames[1:]:
                b = bytearray(f)
                if len(b) > 4:
                    for _ in range(1 + trial % 3):
                        b[int(rng.integers(1, len(b)))] ^= int(1 << rng.integers(0, 8))
                bad.append(bytes(b))
            check(w, h, 24, bad, keys)


@pytest.mark.parametrize("version,bpp", [(2, 24), (3, 24), (4, 24), (2, 16)])
def test_fuzzed_batches_match_the_oracle(version, bpp):
    """tools/sp_fuzz.py in small: corrupted I and P frames (bit flips, truncation), narrow and odd picture sizes, many
    corrupted copies decoded as one batch."""
    rng = np.random.default_rng(4242 + version + bpp)
    for (w, h) in ((33, 17), (16, 16), (9, 20), (96, 32)):
        frames, keys, pics = synth.sp_stream(w, h, 6, seed=w + version, version=version, gop=4, change_permille=200, bpp=bpp)
        specs, cases = [], []
        for trial in range(16):
            bad = []
            for fi, f in enumerate(frames):
                b = bytearray(f)
                hit = (fi > 0) if trial % 4 < 2 else (rng.random() < 0.5)
                if hit and len(b) > 6:
                    for _ in range(1 + trial % 5):
                        b[int(rng.integers(1, len(b)))] ^= int(1 << rng.integers(0, 8))
                    if trial % 4 == 3 and rng.random() < 0.3:
                        b = b[: int(rng.integers(2, len(b)))]
                bad.append(bytes(b))
            specs.append(StreamSpec(SP, w, h, bpp, frames=bad, keys=keys)); cases.append(bad)
        outs, flags = gpu_decode(specs, insign=16)
        k = 0
        for trial, bad in enumerate(cases):
            exp, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, bpp, bad, keys=keys, insignificant_lines=16)
            for i in range(len(bad)):
                err = bool(flags[k] & _lib.JSP_FRAME_ERROR)
                assert err == (st[i] != 0), (w, h, trial, i)
                assert (outs[k] == exp[i]).all(), (w, h, trial, i)
                if not err:
                    assert bool(flags[k] & _lib.JSP_FRAME_CHANGED) == bool(ch[i]), (w, h, trial, i)
                k += 1


@pytest.mark.parametrize("version", [2, 4])
def test_pictures_wider_than_the_shared_memory_ring(version):
    """A picture wider than 16 318 pixels: the last X + 1 pixels no longer fit the I-frame kernels' shared-memory ring (64 KB),
    so the reconstruction warp reads the row above from the picture in HBM.  Same kernels, same model-state layout."""
    w, h = 16400, 24
    frames, keys, pics = synth.sp_stream(w, h, 3, seed=99 + version, version=version, gop=2, change_permille=30)
    check(w, h, 24, frames, keys)
