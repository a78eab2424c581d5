"""Pins the CPU oracle's MSVideo1 restatement (oracle/msvideo1_oracle.c) against
 - the hand-derived known-answer vectors of SURVEY.md Appendix G (from reference src/MSVideo1.hx), and
 - FFmpeg's independent msvideo1 decoder (tests/golden/msv1_ffmpeg_*.npz, made by make_msv1_ffmpeg_golden.py).
The reference itself has no tests or fixtures."""
import numpy as np
import pytest

from conftest import GOLDEN_MSV1, load_golden
import synth
from oracle import pyoracle as O


def rgb15(c):   # MSVideo1.hx:211-214
    return ((c & 0x1F) << 3) + ((c & 0x3E0) << 6) + ((c & 0x7C00) << 9)


def test_kat_g1_onecolour_and_skip():
    f0 = bytes.fromhex("00FCE0831F809EAA")
    out, ch, sg, st = O.decode_stream(O.CODEC_MSVC16, 8, 8, 16, [f0, bytes.fromhex("0484")])
    assert [int(out[0, y, x]) for y, x in [(0, 0), (0, 4), (4, 0), (4, 4)]] == [0xF80000, 0x00F800, 0x0000F8, 0x50A0F0]
    for by in range(2):
        for bx in range(2):
            blk = out[0, by * 4:by * 4 + 4, bx * 4:bx * 4 + 4]
            assert (blk == blk[0, 0]).all()
    # frame 1 is only skips and shorter than size_of_just_skips: previous buffer returned, no significance
    assert list(ch) == [1, 0] and list(sg) == [0, 0]
    assert (out[1] == out[0]).all()


def test_kat_g2_two_colour():
    f = bytes.fromhex("3412007C1F00")
    out, *_ = O.decode_stream(O.CODEC_MSVC16, 4, 4, 16, [f])
    c0, c1 = 0xF80000, 0x0000F8
    exp = [[c1, c1, c0, c1], [c0, c0, c1, c1], [c1, c0, c1, c1], [c0, c1, c1, c1]]
    assert out[0].tolist() == exp


def test_kat_g3_eight_colour():
    flags = 0x5AC3
    cols = [0x8000 | (i * 0x0421) for i in range(1, 9)]
    f = bytes([flags & 0xFF, flags >> 8]) + b"".join(bytes([c & 0xFF, c >> 8]) for c in cols)
    out, *_ = O.decode_stream(O.CODEC_MSVC16, 4, 4, 16, [f])
    for y in range(4):
        for x in range(4):
            bit = ((flags ^ 0xFFFF) >> (4 * y + x)) & 1
            assert int(out[0, y, x]) == rgb15(cols[((y & 2) << 1) + (x & 2) + bit])


def test_kat_g4_8bit_with_skip_carry():
    pal = synth.random_palette(7)
    P = np.frombuffer(pal, dtype="<u4").astype(np.int64)
    f0 = bytes.fromhex("5D2B0A14" "C5B31E1F2021222324254D80C888".replace(" ", ""))
    f1 = bytes.fromhex("028405810184")
    out, ch, sg, st = O.decode_stream(O.CODEC_MSVC8, 8, 8, 8, [f0, f1], palette=pal)
    # block 0: 2 colours, flag bit 1 -> pal[10], 0 -> pal[20]
    fl = 0x2B5D
    for y in range(4):
        for x in range(4):
            assert int(out[0, y, x]) == int(P[10] if (fl >> (4 * y + x)) & 1 else P[20])
    fl8 = 0xB3C5 ^ 0xFFFF
    for y in range(4):
        for x in range(4):
            assert int(out[0, y, 4 + x]) == int(P[30 + ((y & 2) << 1) + (x & 2) + ((fl8 >> (4 * y + x)) & 1)])
    assert (out[0, 4:, :4] == int(P[77])).all() and (out[0, 4:, 4:] == int(P[200])).all()
    # frame 1: blocks 0-1 kept, block 2 = pal[5], block 3 kept
    assert (out[1, :4] == out[0, :4]).all()
    assert (out[1, 4:, :4] == int(P[5])).all() and (out[1, 4:, 4:] == int(P[200])).all()
    assert list(ch) == [1, 1]


@pytest.mark.parametrize("name", GOLDEN_MSV1)
def test_oracle_matches_ffmpeg(name):
    z, frames = load_golden(name)
    is8 = bool(z["is8"])
    out, ch, sg, st = O.decode_stream(O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16, int(z["width"]), int(z["height"]),
                                      8 if is8 else 16, frames, palette=z["palette"].tobytes() if is8 else None)
    assert ((out & int(z["mask"])) == z["expected"]).all()
    if not is8:
        assert ((out & 0x070707) == 0).all()      # fromRGB15 leaves the low 3 bits zero


def test_iskeyframe():
    for is8 in (False, True):
        codec = O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16
        d = O.OracleCodec(codec, 64, 48, 8 if is8 else 16, synth.random_palette(1) if is8 else None)
        assert d.NeedsIndex()
        assert not d.IsKeyFrame(b"")
        assert d.IsKeyFrame(synth.msv1_frame(is8, 64, 48, 5))
        assert not d.IsKeyFrame(synth.msv1_frame(is8, 64, 48, 6, skip_permille=200))


def test_truncated_16bit_reads_undefined():
    """JavaScript semantics on truncated input (MSVideo1.hx:127-181 with out-of-bounds typed-array reads)."""
    full = synth.msv1_frame(False, 32, 16, 3)
    ref, *_ = O.decode_stream(O.CODEC_MSVC16, 32, 16, 16, [full])
    for cut in (len(full) // 2, len(full) // 2 + 1, 7, 1):
        out, ch, *_ = O.decode_stream(O.CODEC_MSVC16, 32, 16, 16, [full[:cut]])
        assert ch[0] == 1
        # the tail is painted with colour 0 one block at a time; everything decoded before the cut is intact
        assert (out[0] != ref[0]).any()
        nb = 0
        for by in range(4):
            for bx in range(8):
                if (out[0, by * 4:by * 4 + 4, bx * 4:bx * 4 + 4] == ref[0, by * 4:by * 4 + 4, bx * 4:bx * 4 + 4]).all():
                    nb += 1
        assert nb >= 1 or cut < 8


def test_significance_16bit():
    w, h = 32, 32
    f0 = synth.msv1_frame(False, w, h, 1)
    # frame 1 changes only block row 0; with 8 insignificant lines (2 block rows) it is not significant
    one = bytes([0x1F, 0x80])
    f1 = one * 8 + bytes([(64 - 8) & 0xFF, 0x84 + ((64 - 8) >> 8)])
    f1 += b"\0" * 40                               # longer than size_of_just_skips so the full path runs
    out, ch, sg, st = O.decode_stream(O.CODEC_MSVC16, w, h, 16, [f0, f1], keys=[0, 0], insignificant_lines=8)
    assert list(ch) == [1, 1] and list(sg) == [1, 0]
    out, ch, sg, st = O.decode_stream(O.CODEC_MSVC16, w, h, 16, [f0, f1], keys=[0, 0], insignificant_lines=0)
    assert list(sg) == [1, 1]


def test_display_convert_kat():
    """Manager.hx:379: 0x00RRGGBB -> 0xFF000000 | B << 16 | G << 8 | R (the Int32 view of canvas bytes R,G,B,A);
    Manager.hx:369: ScreenPressor 16 bpp -> 0xFF000000 | c << 3; flip = Main.hx:946."""
    a = np.array([[0x00112233, 0x00AABBCC], [0x00000001, 0x00FF0000]], dtype=np.int32)
    d = O.display_convert(a).view(np.uint32)
    assert d.tolist() == [[0xFF332211, 0xFFCCBBAA], [0xFF010000, 0xFF0000FF]]
    d = O.display_convert(a, flip=True).view(np.uint32)
    assert d.tolist() == [[0xFF010000, 0xFF0000FF], [0xFF332211, 0xFFCCBBAA]]
    r = np.array([[0x001F1F1F, 0x00010203]], dtype=np.int32)
    assert O.display_convert(r, from_rgb15=True).view(np.uint32).tolist() == [[0xFFF8F8F8, 0xFF081018]]
