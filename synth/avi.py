"""Minimal RIFF/AVI writer for synthetic streams (test/bench input only).

Layout is what the reference's AVIParser walks (reference src/AVIParser.hx:142-184): `avih` main
header (:42-62), one `strl` with `strh` + `strf` (BITMAPINFOHEADER, bpp at offset 14, fourcc at 16,
palette from offset 40 -- :64-88), `LIST movi` of `00dc` chunks padded to even sizes (:144), and an
`idx1` index whose flag 0x10 marks key frames with offsets relative to the `movi` fourcc
(src/DataLoaderAVIIndexed.hx:298-323).
"""
import struct


def _chunk(fourcc, payload):
    pad = b"\0" if len(payload) & 1 else b""
    return fourcc + struct.pack("<I", len(payload)) + payload + pad


def _list(kind, payload):
    return b"LIST" + struct.pack("<I", len(payload) + 4) + kind + payload


def avi_bytes(width, height, bpp, fourcc, frames, keys=None, palette=None, fps=15, top_down=False):
    n = len(frames)
    if keys is None:
        keys = [1] + [0] * (n - 1)
    usec = int(1000000 / fps)
    max_len = max([len(f) for f in frames] + [0])
    avih = struct.pack("<IIIIIIIIII4I", usec, max_len * fps, 0, 0x10, n, 0, 1, max_len, width, height, 0, 0, 0, 0)
    strh = struct.pack("<4s4sIHHIIIIIIII4h", b"vids", fourcc, 0, 0, 0, 0, 1, fps, 0, n, max_len, 0xFFFFFFFF, 0,
                       0, 0, width, height)
    ncol = 256 if bpp == 8 else 0
    bih = struct.pack("<IiiHH4sIiiII", 40, width, -height if top_down else height, 1, bpp, fourcc,
                      width * height * (bpp // 8), 0, 0, ncol, 0)
    strf = bih + (palette if (bpp == 8 and palette) else b"")
    hdrl = _list(b"hdrl", _chunk(b"avih", avih) + _list(b"strl", _chunk(b"strh", strh) + _chunk(b"strf", strf)))
    movi_payload = b""
    idx = b""
    off = 4                                   # offsets in idx1 are relative to the 'movi' fourcc
    for f, k in zip(frames, keys):
        c = _chunk(b"00dc", bytes(f))
        idx += struct.pack("<4sIII", b"00dc", 0x10 if k else 0, off, len(f))
        off += len(c)
        movi_payload += c
    movi = _list(b"movi", movi_payload)
    body = b"AVI " + hdrl + movi + _chunk(b"idx1", idx)
    return b"RIFF" + struct.pack("<I", len(body)) + body


def write_avi(path, *args, **kw):
    with open(path, "wb") as fh:
        fh.write(avi_bytes(*args, **kw))
