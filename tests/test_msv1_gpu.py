"""GPU parity: MSVideo1 decoded by the CUDA path through the C ABI must be bit-exact against the CPU oracle
(and against FFmpeg's golden pictures) -- pictures, `changed` and `significant_changes`."""
import numpy as np
import pytest

from conftest import GOLDEN_MSV1, load_golden
from jsplayer_b200 import BatchDecoder, StreamSpec, MSVideo1_16bit, MSVideo1_8bit, CodecType, DecoderState
import synth
from jsplayer_b200 import _lib
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def oracle_stream(is8, w, h, frames, keys=None, pal=None, insign=0):
    return O.decode_stream(O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16, w, h, 8 if is8 else 16, frames, keys=keys,
                           palette=pal, insignificant_lines=insign)


def gpu_streams(specs, insign=0, significance=True):
    bd = BatchDecoder(insignificant_lines=insign, significance=significance)
    bd.configure(specs)
    outs, flags = bd.decode_host()
    bd.close()
    return outs, flags


def check_stream(is8, w, h, frames, keys=None, insign=0, pal_seed=1):
    pal = synth.random_palette(pal_seed) if is8 else None
    exp, ch, sg, st = oracle_stream(is8, w, h, frames, keys, pal, insign)
    spec = StreamSpec(CodecType.codec_msvc8 if is8 else CodecType.codec_msvc16, w, h, 8 if is8 else 16,
                      frames=frames, keys=keys, palette=pal)
    outs, flags = gpu_streams([spec], insign)
    bw, bh = w & ~3, h & ~3
    for i in range(len(frames)):
        assert (outs[i][:bh, :bw] == exp[i][:bh, :bw]).all(), "frame %d differs" % i
        assert bool(flags[i] & _lib.JSP_FRAME_CHANGED) == bool(ch[i]), "changed flag of frame %d" % i
        if not (keys is not None and keys[i]) and not (keys is None and i == 0):
            assert bool(flags[i] & _lib.JSP_FRAME_SIGNIFICANT) == bool(sg[i]), "significant flag of frame %d" % i
        assert not (flags[i] & _lib.JSP_FRAME_ERROR)


@pytest.mark.parametrize("is8", [False, True])
@pytest.mark.parametrize("size", [(4, 4), (8, 8), (64, 48), (320, 240), (132, 100), (1920, 1080)])
def test_key_frames(is8, size):
    w, h = size
    for seed, mix in [(1, (25, 50, 25)), (2, (100, 0, 0)), (3, (0, 100, 0)), (4, (0, 0, 100))]:
        check_stream(is8, w, h, [synth.msv1_frame(is8, w, h, seed * 7 + w, mix=mix)])


@pytest.mark.parametrize("is8", [False, True])
def test_p_frames_with_skips(is8):
    w, h = 320, 240
    frames = [synth.msv1_frame(is8, w, h, 50)]
    for i in range(1, 12):
        frames.append(synth.msv1_frame(is8, w, h, 50 + i, skip_permille=[20, 200, 900][i % 3], mean_skip=[3, 40, 400][i % 3],
                                       mix=(40, 40, 20)))
    check_stream(is8, w, h, frames, insign=36)
    check_stream(is8, w, h, frames, insign=0)


@pytest.mark.parametrize("is8", [False, True])
def test_unchanged_and_empty_frames(is8):
    w, h = 64, 48
    nb = (w // 4) * (h // 4)
    key = synth.msv1_frame(is8, w, h, 9)
    skip_all = bytes([nb & 0xFF, 0x84 + (nb >> 8)])
    rest = bytes([0x00, 0x84])                                # skip count 0: rest of the frame is copied
    frames = [key, b"", skip_all, rest + b"\x1f\x80" * 30, key[: len(key) // 2], skip_all + b"\0" * 64, key]
    check_stream(is8, w, h, frames, keys=[1, 0, 0, 0, 0, 0, 0], insign=8)


@pytest.mark.parametrize("is8", [False, True])
def test_truncated_frames(is8):
    w, h = 64, 32
    full = synth.msv1_frame(is8, w, h, 77, mix=(20, 40, 40))
    cuts = sorted(set([1, 2, 3, 5, 6, 7, 9, 17, 18, 19, len(full) // 2, len(full) // 2 + 1, len(full) - 1, len(full) - 2,
                       len(full) - 3]))
    frames = [full] + [full[:c] for c in cuts]
    check_stream(is8, w, h, frames, keys=[1] + [0] * len(cuts))


def test_8bit_terminator():
    w, h = 32, 16
    key = synth.msv1_frame(True, w, h, 5)
    f = b"\x07\x80" * 5 + b"\0\0" + b"\x09\x80" * 40
    check_stream(True, w, h, [key, f, f[:10], b"\0\0"], keys=[1, 0, 0, 0])


@pytest.mark.parametrize("is8", [False, True])
def test_trailing_garbage_and_pad_byte(is8):
    w, h = 64, 64
    f = synth.msv1_frame(is8, w, h, 21)
    rng = np.random.default_rng(1)
    frames = [f + b"\0", f + rng.integers(0, 256, 301, dtype=np.uint8).tobytes(), f + b"\x00\x84"]
    check_stream(is8, w, h, frames, keys=[1, 1, 1])


@pytest.mark.parametrize("is8", [False, True])
def test_bytes_after_a_terminator_hold_a_second_terminator(is8):
    """Skip count 0 (`00 84`) / the 8-bit `00 00` end the frame: the rest is copied from the previous picture.  Lanes
    past the terminator still parse the trailing bytes; a SECOND terminator pattern among them must not move the
    "rest of frame" copy (sm.big_blk0 used to be a plain store racing with the real one)."""
    w, h = 256, 128
    rng = np.random.default_rng(7)
    key = synth.msv1_frame(is8, w, h, 3)
    term = b"\x00\x00" if is8 else b"\x00\x84"
    frames, keys = [key], [1]
    for trial in range(12):
        head = synth.msv1_frame(is8, w, h, 50 + trial, skip_permille=300)[: int(rng.integers(2, 600)) & ~1]
        tail = rng.integers(0, 256, size=int(rng.integers(4, 900)) & ~1, dtype=np.uint8).tobytes()
        junk = b"".join(tail[i:i + 32] + term for i in range(0, len(tail), 32))      # many more terminator patterns
        frames.append(head + term + junk); keys.append(0)
    check_stream(is8, w, h, frames, keys=keys)


def test_random_bytes_do_not_crash_and_match():
    """Arbitrary bytes are a valid opcode stream for this codec; the walk must agree with the oracle."""
    rng = np.random.default_rng(99)
    w, h = 128, 96
    for is8 in (False, True):
        frames = [synth.msv1_frame(is8, w, h, 1)]
        frames += [rng.integers(0, 256, size=int(n), dtype=np.uint8).tobytes() for n in (3000, 9000, 20000, 12001, 40)]
        check_stream(is8, w, h, frames, keys=[1] + [0] * 5)


def test_many_streams_mixed_sizes():
    specs, exp = [], []
    for s in range(24):
        is8 = bool(s & 1)
        w, h = [(64, 48), (320, 240), (100, 60), (640, 360)][s % 4]
        pal = synth.random_palette(s) if is8 else None
        frames = [synth.msv1_frame(is8, w, h, 1000 + s)] + [
            synth.msv1_frame(is8, w, h, 2000 + s * 10 + i, skip_permille=100, mean_skip=20) for i in range(s % 5)]
        keys = [1] + [0] * (len(frames) - 1)
        specs.append(StreamSpec(CodecType.codec_msvc8 if is8 else CodecType.codec_msvc16, w, h, 8 if is8 else 16,
                                frames=frames, keys=keys, palette=pal))
        exp.append(oracle_stream(is8, w, h, frames, keys, pal)[0])
    outs, flags = gpu_streams(specs)
    i = 0
    for s, e in enumerate(exp):
        for f in range(e.shape[0]):
            bh, bw = e.shape[1] & ~3, e.shape[2] & ~3
            assert (outs[i][:bh, :bw] == e[f][:bh, :bw]).all(), (s, f)
            i += 1


def test_key_flag_lies_are_repaired():
    """A frame flagged as key that still copies from its predecessor is re-decoded in order."""
    w, h = 64, 48
    frames = [synth.msv1_frame(False, w, h, 1), synth.msv1_frame(False, w, h, 2, skip_permille=300),
              synth.msv1_frame(False, w, h, 3, skip_permille=300)]
    exp, *_ = oracle_stream(False, w, h, frames, keys=[1, 1, 1])
    outs, flags = gpu_streams([StreamSpec(CodecType.codec_msvc16, w, h, 16, frames=frames, keys=[1, 1, 1])])
    for i in range(3):
        assert (outs[i] == exp[i]).all()
        assert not (flags[i] & _lib.JSP_FRAME_ERROR)


@pytest.mark.parametrize("name", GOLDEN_MSV1)
def test_gpu_matches_ffmpeg_golden(name):
    z, frames = load_golden(name)
    is8 = bool(z["is8"])
    w, h = int(z["width"]), int(z["height"])
    spec = StreamSpec(CodecType.codec_msvc8 if is8 else CodecType.codec_msvc16, w, h, 8 if is8 else 16, frames=frames,
                      palette=z["palette"].tobytes() if is8 else None)
    outs, flags = gpu_streams([spec])
    for i in range(len(frames)):
        assert ((outs[i] & int(z["mask"])) == z["expected"][i]).all()


@pytest.mark.parametrize("is8", [False, True])
def test_per_stream_dropin_matches_oracle(is8):
    """IVideoCodec members through the C ABI, driven the way Manager.worker drives them (Manager.hx:454-525)."""
    w, h = 160, 120
    pal = synth.random_palette(4) if is8 else None
    mine = MSVideo1_8bit(w, h, pal) if is8 else MSVideo1_16bit(w, h)
    ora = O.OracleCodec(O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16, w, h, 8 if is8 else 16, pal)
    mine.Preinit(36); ora.Preinit(36)
    nb = (w // 4) * (h // 4)
    frames = [synth.msv1_frame(is8, w, h, 31)]
    frames += [synth.msv1_frame(is8, w, h, 32 + i, skip_permille=150, mean_skip=30) for i in range(5)]
    frames += [b"", bytes([nb & 0xFF, 0x84 + (nb >> 8)]), synth.msv1_frame(is8, w, h, 60, skip_permille=950, mean_skip=900)]
    bufs_m = [np.zeros(w * h, dtype=np.int32) for _ in range(3)]
    bufs_o = [np.zeros(w * h, dtype=np.int32) for _ in range(3)]
    assert mine.PreviousFrame() is None
    for i, f in enumerate(frames):
        km, ko = mine.IsKeyFrame(f), ora.IsKeyFrame(f)
        assert km == ko
        dm = next(b for b in bufs_m if b is not mine.PreviousFrame())
        do = next(b for b in bufs_o if b is not ora.PreviousFrame())
        if i == 0:
            assert mine.DecompressI(f, dm) == DecoderState.zero_state
            ora.DecompressI(f, do)
            pm, po = mine.PreviousFrame(), ora.PreviousFrame()
            sm = so = False
        else:
            r = mine.DecompressP(f, dm)
            po, so = ora.DecompressP(f, do)
            pm, sm = r.data_pnt, r.significant_changes
        assert (pm is dm) == (po is do), "frame %d: data_pnt identity" % i
        assert sm == so, "frame %d: significant_changes" % i
        assert pm is not None and (pm == po).all(), "frame %d: picture" % i
    mine.StopAndClean()
    assert mine.PreviousFrame() is None


def test_full_size_properties_1080p_batch():
    """BASELINE config 2 shape at reduced count: property checks that need no oracle pass over every pixel."""
    w, h, n = 1920, 1080, 8
    specs = [StreamSpec(CodecType.codec_msvc16, w, h, 16, frames=[synth.msv1_frame(False, w, h, 500 + i, mix=(25, 50, 25))])
             for i in range(n)]
    outs, flags = gpu_streams(specs)
    for i in range(n):
        o = outs[i]
        assert ((o.view(np.uint32) & np.uint32(0xFF070707)) == 0).all()                   # RGB555 -> 0x00RRGGBB with low 3 bits clear
        assert flags[i] & _lib.JSP_FRAME_CHANGED
    # identical input -> identical output (idempotence), and one oracle spot check
    outs2, _ = gpu_streams(specs[:2])
    assert (outs2[0] == outs[0]).all() and (outs2[1] == outs[1]).all()
    exp, *_ = oracle_stream(False, w, h, list(specs[3].frames))
    assert (outs[3] == exp[0]).all()


def test_persistent_mode_large_batch():
    """Enough 4 KiB bitstream tiles (> 6 x 8 CTAs x SM count) to put msv1_decode_kernel into its persistent,
    prefetching mode: key + P frames with skip runs at 1080p, RGB555 and 8-bit, against the oracle."""
    w, h = 1920, 1080
    specs, exp = [], []
    pal = synth.random_palette(5)
    for s in range(64):
        is8 = s % 2 == 1
        frames = [synth.msv1_frame(is8, w, h, 1000 + s, mix=(10, 30, 60))] + \
                 [synth.msv1_frame(is8, w, h, 2000 + s, skip_permille=300, mean_skip=25, mix=(10, 30, 60))]
        specs.append(StreamSpec(CodecType.codec_msvc8 if is8 else CodecType.codec_msvc16, w, h, 8 if is8 else 16,
                                frames=frames, palette=pal if is8 else None))
        if s % 16 in (0, 1, 15):
            exp.append((2 * s, O.decode_stream(O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16, w, h, 8 if is8 else 16, frames,
                                               palette=pal if is8 else None)[0]))
    bd = BatchDecoder()
    bd.configure(specs)
    outs, flags = bd.decode_host()
    bd.close()
    assert not (flags & _lib.JSP_FRAME_ERROR).any()
    for first, e in exp:
        for f in range(2):
            assert (outs[first + f] == e[f]).all(), "frame %d" % (first + f)


def test_absurd_picture_sizes_are_refused():
    """A corrupt AVI header can claim any 32-bit size: the batcher refuses it instead of sizing tables from it."""
    frame = synth.msv1_frame(False, 64, 48, 1)
    for (w, h) in ((1 << 20, 1 << 20), (40000, 16), (16, 40000), (32768, 16384)):
        bd = BatchDecoder()
        with pytest.raises(RuntimeError, match="out of range"):
            bd.configure([StreamSpec(CodecType.codec_msvc16, w, h, 16, frames=[frame], keys=[1])])
        bd.close()


@pytest.mark.parametrize("height,mix", [(3112, (0, 0, 100)), (3200, (0, 0, 100)), (3112, (25, 50, 25))])
def test_one_frame_whose_lookback_chain_exceeds_residency(height, mix):
    """ONE frame, ~7000 bitstream tiles in a single look-back chain.  8192 x 3112 with 18-byte blocks is 28.7 MB = 7002 tiles:
    below the persistent-grid threshold (6 x 8 x 148 = 7104), so the launch is 7002 one-shot CTAs for 1184 resident slots --
    the chain crosses several waves and forward progress rests on the tile-major tickets alone.  8192 x 3200 (7200 tiles)
    takes the persistent grid.  The look-back watchdog (MSV1_SPIN_LIMIT) would fail the frame instead of hanging: the
    frame must decode, error-free and bit-exact, twice in a row (the tile states are reset between runs)."""
    w = 8192
    frame = synth.msv1_frame(False, w, height, 0x7000 + height, mix=mix)
    if mix == (0, 0, 100):
        # the one-shot / persistent threshold is 6 x resident CTAs x tile bytes = 29 097 984 bytes for either tile size
        assert (28_600_000 <= len(frame) < 29_097_984) if height == 3112 else (len(frame) >= 29_097_984)
    exp = oracle_stream(False, w, height, [frame])[0]
    bd = BatchDecoder()
    bd.configure([StreamSpec(CodecType.codec_msvc16, w, height, 16, frames=[frame])])
    for _ in range(2):
        outs, flags = bd.decode_host()
        assert not (flags & _lib.JSP_FRAME_ERROR).any()
        assert (outs[0] == exp[0]).all()
    bd.close()
