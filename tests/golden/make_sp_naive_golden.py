"""Generates tests/golden/sp_naive_*.npz: synthetic ScreenPressor streams together with the pictures AND the per-symbol
(call kind, symbol, freq, cumFreq, total) traces that oracle/sp_naive.py -- the second, independent, line-by-line Python
reading of ANS.hx / RangeCoder.hx / EntroCoders.hx / ScreenPressor.hx -- decodes from them.

    python tests/golden/make_sp_naive_golden.py            # rewrites the fixtures (about a minute)

The fixtures pin the C oracle (tests/test_sp_second_reading.py: oracle pictures == these, oracle symbol trace == these)
and the CUDA path (-m gpu: pictures through the C ABI == these).  The streams come from the synthetic encoder (synth/);
what is pinned is the DECODER: two readings of the reference's text, written at different times without sight of each
other, agree symbol by symbol.  Each case also records which rare model paths it exercised (`events`), and the generator
refuses to write a corpus that does not cover every context-kind transition, every rescale/rebuild path and both coders.
It also decodes every stream a second time with the P-frame destination buffers pre-filled with a junk value instead of
the previous picture, and asserts the pictures do not change: nothing in the corpus depends on what a stale ring buffer
holds (ScreenPressor.hx:440-449 at a block's first column), which is the one place where this repository defines
behaviour the reference leaves to chance.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import synth                                         # noqa: E402
from oracle import sp_naive as N                     # noqa: E402


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


def case_screen(version, w, h, n, gop, bpp=24, seed=1, cp=60):
    frames, keys, _ = synth.sp_stream(w, h, n, seed=seed, version=version, gop=gop, change_permille=cp, bpp=bpp)
    return w, h, bpp, frames, keys


def case_allkinds(version, w=192, h=128):
    enc = synth.SPEncoder(w, h, 24, version)
    px = many_symbol_picture(w, h, 5)
    f0 = enc.iframe(px)
    nxt = synth.noise(w, h, 6)
    nxt[: h // 2] = px[: h // 2]
    return w, h, 24, [f0, enc.pframe(nxt, px)], [1, 0]


def case_flat(version, w=64, h=32):
    """coded I, P, empty frame, `unchanged` P, flat key frame (models reset), P on the flat picture, coded I, P."""
    enc = synth.SPEncoder(w, h, 24, version)
    p0 = synth.screen(w, h, 1)
    p1, mv = synth.screen_next(p0, 2, 100)
    f_i = enc.iframe(p0)
    f_p = enc.pframe(p1, p0, mv)
    f_flat = enc.flat(0x123456)
    flat_pic = np.full((h, w), 0x123456, dtype=np.int32)
    p2, mv2 = synth.screen_next(flat_pic, 3, 100)
    f_p2 = enc.pframe(p2, flat_pic, mv2)
    p3 = synth.screen(w, h, 9)
    f_i2 = enc.iframe(p3)
    p4, mv4 = synth.screen_next(p3, 4, 150)
    f_p4 = enc.pframe(p4, p3, mv4)
    return w, h, 24, [f_i, f_p, b"", b"\0", f_flat, f_p2, f_i2, f_p4], [1, 0, 0, 0, 1, 0, 1, 0]


def case_noise_v2(w=128, h=96):
    enc = synth.SPEncoder(w, h, 24, 2)
    px = synth.noise(w, h, 21, ncolors=(2, 3, 5, 17))
    f0 = enc.iframe(px)
    nxt = synth.noise(w, h, 22, ncolors=(2, 3, 5, 17))
    nxt[: h // 3] = px[: h // 3]
    return w, h, 24, [f0, enc.pframe(nxt, px)], [1, 0]


CASES = {
    "v2_screen_96x64": lambda: case_screen(2, 96, 64, 8, 4),
    "v3_screen_96x64": lambda: case_screen(3, 96, 64, 8, 4),
    "v4_screen_96x64": lambda: case_screen(4, 96, 64, 8, 4),
    "v2_screen_33x17": lambda: case_screen(2, 33, 17, 5, 0, seed=7, cp=150),
    "v4_screen_33x17": lambda: case_screen(4, 33, 17, 5, 0, seed=7, cp=150),
    "v2_16bpp_80x48": lambda: case_screen(2, 80, 48, 6, 3, bpp=16, seed=5),
    "v3_16bpp_80x48": lambda: case_screen(3, 80, 48, 6, 3, bpp=16, seed=5),
    "v2_noise_128x96": case_noise_v2,
    "v3_allkinds_192x128": lambda: case_allkinds(3),
    "v4_allkinds_192x128": lambda: case_allkinds(4),
    "v2_flat_64x32": lambda: case_flat(2),
    "v4_flat_64x32": lambda: case_flat(4),
}

REQUIRED_EVENTS = ["kind_0_to_1", "kind_1_to_4", "kind_1_to_5", "kind_1_to_2", "kind_4_to_5", "kind_5_to_6", "kind_2_to_6",
                   "kind_2_to_3", "kind_3_to_7", "kind_6_to_7", "cx6_grow", "cx6_rescaleDec", "small_rescale_S4",
                   "small_rescale_S16", "fixed_rebuild_256", "fixed_rebuild_6", "rc_rescale_uni", "rc_rescale_256",
                   "rc_rescale_6"]


def main():
    total = 0
    all_events = {}
    for name, make in CASES.items():
        w, h, bpp, frames, keys = make()
        N.EVENTS.clear()
        pics, changed, signif, traces, info = N.decode_stream(w, h, bpp, frames, trace=True)
        events = dict(N.EVENTS)
        pics2, changed2, signif2, _, _ = N.decode_stream(w, h, bpp, frames, stale=0x00ABCDEF)
        assert pics2 == pics and changed2 == changed and signif2 == signif, \
            "%s: the pictures depend on stale destination-buffer contents" % name
        assert info.get("oob_reads", 0) == 0, "%s: the rANS reader ran past the end of a frame" % name
        n = len(frames)
        out = np.zeros((n, h, w), dtype=np.int32)
        have = np.zeros(n, dtype=np.uint8)
        last = None
        for f in range(n):
            if pics[f] is not None:
                last = np.array(pics[f], dtype=np.int64).astype(np.int32).reshape(h, w)
                have[f] = 1
            if last is not None:
                out[f] = last
        tr = np.array([t for fr in traces for t in fr], dtype=np.int32).reshape(-1, 5)
        tr_off = np.cumsum([0] + [len(fr) for fr in traces]).astype(np.int64)
        path = os.path.join(HERE, "sp_naive_%s.npz" % name)
        np.savez_compressed(path, data=np.frombuffer(b"".join(frames), dtype=np.uint8),
                            frame_len=np.array([len(f) for f in frames], dtype=np.uint32),
                            keys=np.array(keys, dtype=np.uint8), width=w, height=h, bpp=bpp,
                            pictures=out, have_picture=have, changed=np.array(changed, dtype=np.uint8),
                            significant=np.array(signif, dtype=np.uint8), trace=tr, trace_frame_off=tr_off,
                            events=json.dumps(events, sort_keys=True))
        sz = os.path.getsize(path)
        total += sz
        for k, v in events.items():
            all_events[k] = all_events.get(k, 0) + v
        print("%-24s %4dx%-4d bpp %2d  %d frames  %7d symbols  %7d B  %s" % (name, w, h, bpp, n, tr.shape[0], sz, events))
    missing = [k for k in REQUIRED_EVENTS if all_events.get(k, 0) == 0]
    assert not missing, "the corpus never exercised: %s" % missing
    print("total %d bytes; every required model path exercised: %s" % (total, {k: all_events[k] for k in REQUIRED_EVENTS}))


if __name__ == "__main__":
    main()
