"""The headless mirror of Manager.worker (jsplayer_b200/manager.py) on the CPU: driven over the oracle's IVideoCodec it must
show the right picture for every request of a script of plays, seeks and still-skips, restart from key frames, never hand
out the buffer the codec still borrows, and find stills the way DataLoader.FindPossibleChange defines them.  (The same script
runs over the GPU drop-in in tests/test_manager_gpu.py and is compared step by step.)"""
import importlib.util
import os

import numpy as np
import pytest

from jsplayer_b200.manager import Manager
from oracle import pyoracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("_mgr_gpu", os.path.join(HERE, "test_manager_gpu.py"))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)


@pytest.mark.parametrize("kind", ["sp2", "sp4", "msv16", "msv8"])
def test_manager_mirror_over_the_oracle_codec(kind):
    w, h, bpp, pal, frames, keys = _m.make_stream(kind)
    codec = O.CODEC_SCREENPRESSOR if kind.startswith("sp") else (O.CODEC_MSVC8 if kind == "msv8" else O.CODEC_MSVC16)
    dec = O.OracleCodec(codec, w, h, bpp, palette=pal) if pal else O.OracleCodec(codec, w, h, bpp)
    m = Manager(dec, w, h, frames, keys, nbuffers=9)
    exp, ch, sg, st = O.decode_stream(codec, w, h, bpp, frames, keys=keys, palette=pal, insignificant_lines=36)
    for step in _m.SCRIPT:
        if step[0] == "show":
            pic = m.show(step[1])
            assert pic is not None and (pic.reshape(h, w) == exp[step[1]]).all(), step
            prev = dec.PreviousFrame()
            assert prev is None or any(prev is b for b in m.buffers)             # the codec borrows one of OUR buffers
        else:
            before = m.frame_of_interest
            pos = m.SkipStills()
            assert pos is not None and pos > before
            # every frame skipped over was known to be a still; the one found is a change (or the last frame)
            for f in range(before + 1, pos):
                assert m.frames[f].significant_changes is False
            assert m.frames[pos].significant_changes or pos == len(frames) - 1
    # P frames: the stored flag is the codec's significant_changes; unchanged frames extended a buffer's range
    for f in range(len(frames)):
        if not keys[f] and m.frames[f].significant_changes is not None:
            assert m.frames[f].significant_changes == bool(sg[f]), f
    assert any(b is not None and b[1] > b[0] for b in m.bufs) or any(b is None for b in m.bufs)
    seeks = sum(1 for i in range(1, len(m.decoded_log)) if m.decoded_log[i][1] < m.decoded_log[i - 1][1])
    assert seeks >= 2


def test_get_free_buffer_never_returns_the_borrowed_buffer():
    w, h, bpp, pal, frames, keys = _m.make_stream("msv16")
    dec = O.OracleCodec(O.CODEC_MSVC16, w, h, 16)
    m = Manager(dec, w, h, frames, keys, nbuffers=3)                             # a tight ring: 3 buffers
    m.frame_of_interest = 10 ** 6                                                # everything decoded is "behind" the player
    for _ in range(len(frames)):
        prev = dec.PreviousFrame()
        prev_idx = m._index_of(prev)
        assert m.worker()
        new_prev = dec.PreviousFrame()
        if prev is not None and new_prev is not prev:
            assert new_prev is not None and m._index_of(new_prev) != prev_idx     # the new picture went to a different buffer
    assert m.next_frame_to_decode == len(frames)
