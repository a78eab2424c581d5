"""The reference's frame scheduler (Manager.worker, the 9-buffer ring, seek-to-key-frame restart, SkipStills) driven over the
per-stream C ABI on the GPU and over the CPU oracle's IVideoCodec: same script, same buffers, same bookkeeping
(jsplayer_b200/manager.py mirrors Manager.hx:114-118, 244-249, 289-317, 392-443, 454-539, 568-578)."""
import numpy as np
import pytest

from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
from jsplayer_b200 import MSVideo1_16bit, MSVideo1_8bit, ScreenPressor
from jsplayer_b200.manager import Manager
import synth
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def make_stream(kind):
    w, h = 96, 64
    if kind == "sp2" or kind == "sp4":
        version = 2 if kind == "sp2" else 4
        enc = synth.SPEncoder(w, h, 24, version)
        cur = synth.screen(w, h, 5)
        frames, keys = [enc.iframe(cur)], [1]
        for f in range(1, 40):
            if f % 10 == 0:
                cur = synth.screen(w, h, 5 + f) if f != 20 else cur      # frame 20: a key frame identical to the picture before it
                frames.append(enc.iframe(cur)); keys.append(1)
            elif f % 3 == 0:
                frames.append(enc.pframe(cur, cur)); keys.append(0)      # a P frame that changes nothing (a "still")
            else:
                nxt, mv = synth.screen_next(cur, 100 + f, 60)
                frames.append(enc.pframe(nxt, cur, mv)); keys.append(0)
                cur = nxt
        return w, h, 24, None, frames, keys
    is8 = kind == "msv8"
    pal = synth.random_palette(3) if is8 else None
    frames, keys = [], []
    nb = (w // 4) * (h // 4)
    for f in range(40):
        if f % 8 == 0:
            frames.append(synth.msv1_frame(is8, w, h, 700 + f)); keys.append(1)
        elif f % 3 == 0:
            frames.append(bytes([nb & 0xFF, 0x84 + (nb >> 8)])); keys.append(0)   # every block skipped: unchanged frame
        else:
            frames.append(synth.msv1_frame(is8, w, h, 700 + f, skip_permille=600, mean_skip=20)); keys.append(0)
    return w, h, 8 if is8 else 16, pal, frames, keys


def make_decoders(kind, w, h, bpp, pal):
    if kind.startswith("sp"):
        return ScreenPressor(w, h, bpp), O.OracleCodec(O.CODEC_SCREENPRESSOR, w, h, bpp)
    if kind == "msv8":
        return MSVideo1_8bit(w, h, pal), O.OracleCodec(O.CODEC_MSVC8, w, h, 8, palette=pal)
    return MSVideo1_16bit(w, h), O.OracleCodec(O.CODEC_MSVC16, w, h, 16)


SCRIPT = [("show", 0), ("show", 1), ("show", 7), ("show", 8), ("show", 19), ("show", 3),      # play, then seek backwards
          ("show", 33), ("show", 34), ("show", 12), ("skip",), ("skip",), ("skip",), ("show", 39), ("show", 38), ("skip",)]


@pytest.mark.parametrize("kind", ["sp2", "sp4", "msv16", "msv8"])
def test_manager_worker_mirror_gpu_equals_cpu(kind):
    w, h, bpp, pal, frames, keys = make_stream(kind)
    gpu_dec, cpu_dec = make_decoders(kind, w, h, bpp, pal)
    mg = Manager(gpu_dec, w, h, frames, keys, nbuffers=9)
    mc = Manager(cpu_dec, w, h, frames, keys, nbuffers=9)
    exp = O.decode_stream(O.CODEC_SCREENPRESSOR if kind.startswith("sp") else (O.CODEC_MSVC8 if kind == "msv8" else O.CODEC_MSVC16),
                          w, h, bpp, frames, keys=keys, palette=pal, insignificant_lines=36)[0]
    for step in SCRIPT:
        if step[0] == "show":
            a, b = mg.show(step[1]), mc.show(step[1])
            assert a is not None and b is not None, step
            assert (a == b).all(), step
            assert (a.reshape(h, w) == exp[step[1]]).all(), step       # and it is the picture of that frame
        else:
            ra, rb = mg.SkipStills(), mc.SkipStills()
            assert ra == rb, (step, ra, rb)
        assert mg.bufs == mc.bufs, step                                 # same ring bookkeeping after every step
        assert mg.next_frame_to_decode == mc.next_frame_to_decode and mg.frame_of_interest == mc.frame_of_interest
        assert mg.decoded_log == mc.decoded_log
        assert [f.significant_changes for f in mg.frames] == [f.significant_changes for f in mc.frames], step
    assert any(b is not None and b[1] > b[0] for b in mg.bufs) or kind            # an unchanged frame extended a buffer's range
    seeks = sum(1 for i in range(1, len(mg.decoded_log)) if mg.decoded_log[i][1] < mg.decoded_log[i - 1][1])
    assert seeks >= 2                                                    # the script really restarted from key frames


@pytest.mark.parametrize("kind", ["sp4", "msv16"])
def test_batch_next_significant_equals_skip_stills(kind):
    """jsp_batch_next_significant over a decoded batch answers what Manager.SkipStills finds by decoding forward."""
    w, h, bpp, pal, frames, keys = make_stream(kind)
    _, cpu_dec = make_decoders(kind, w, h, bpp, pal)
    m = Manager(cpu_dec, w, h, frames, keys, nbuffers=9)
    m.show(0)
    codec = CodecType.codec_screenpressor if kind.startswith("sp") else CodecType.codec_msvc16
    bd = BatchDecoder(insignificant_lines=36, significance=True)
    bd.configure([StreamSpec(codec, w, h, bpp, frames=frames, keys=keys, palette=pal)])
    bd.decode_host()
    pos = 0
    for _ in range(12):
        want = m.SkipStills()
        got = bd.next_significant(0, pos + 1)
        assert got == want, (pos, got, want)
        pos = got
        if pos >= len(frames) - 1:
            break
    bd.close()
