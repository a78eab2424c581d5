"""Generates tests/golden/msv1_ffmpeg_*.npz: synthetic MSVideo1 (CRAM) streams together with the pictures an
INDEPENDENT decoder -- FFmpeg's msvideo1, reached through cv2.VideoCapture -- produces for them.

The reference (Haxe->JS) ships no golden vectors and cannot run here; these fixtures pin the CPU oracle
(oracle/msvideo1_oracle.c) and, through it, the CUDA path.  Run from the repo root:
    python tests/golden/make_msv1_ffmpeg_golden.py
Normalisation of FFmpeg's output to the reference's frame layout:
  * FFmpeg returns the image top-down; the reference keeps bitstream (bottom-up DIB) row order -> flip rows;
  * FFmpeg returns B,G,R bytes; the reference packs 0x00RRGGBB;
  * RGB555: FFmpeg expands 5->8 bits with bit replication, the reference uses <<3 -> compare & 0xF8F8F8
    (the fixture stores FFmpeg's pixels masked that way, and `mask` says so).
"""
import os
import sys
import tempfile

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import synth                      # noqa: E402
from jsplayer_b200.synth.avi import write_avi        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ffmpeg_decode(path, n, w, h):
    cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
    frames = []
    for _ in range(n):
        ok, img = cap.read()
        if not ok:
            break
        img = img[::-1].astype(np.uint32)                           # top-down -> bitstream row order
        frames.append(((img[..., 2] << 16) | (img[..., 1] << 8) | img[..., 0]).astype(np.int32))
    cap.release()
    assert len(frames) == n, "FFmpeg decoded %d of %d frames" % (len(frames), n)
    return np.stack(frames)


def make(name, is8, w, h, n, seed, skip_permille, mix):
    pal = synth.random_palette(seed) if is8 else None
    frames = [synth.msv1_frame(is8, w, h, seed * 1000 + i, skip_permille=0 if i == 0 else skip_permille, mix=mix)
              for i in range(n)]
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, name + ".avi")
        write_avi(p, w, h, 8 if is8 else 16, b"CRAM", frames, palette=pal)
        exp = ffmpeg_decode(p, n, w, h)
    mask = 0xFFFFFF if is8 else 0xF8F8F8
    exp &= mask
    ln = np.array([len(f) for f in frames], dtype=np.uint32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), is8=is8, width=w, height=h, mask=mask,
                        frame_len=ln, data=np.frombuffer(b"".join(frames), dtype=np.uint8),
                        palette=np.frombuffer(pal, dtype=np.uint8) if pal else np.zeros(0, np.uint8),
                        expected=exp)
    print(name, "frames", n, "bytes", int(ln.sum()))


if __name__ == "__main__":
    make("msv1_ffmpeg_rgb555_64x48", False, 64, 48, 6, 11, 300, (25, 50, 25))
    make("msv1_ffmpeg_rgb555_320x240", False, 320, 240, 4, 12, 150, (30, 40, 30))
    make("msv1_ffmpeg_pal8_64x48", True, 64, 48, 6, 13, 300, (25, 50, 25))
    make("msv1_ffmpeg_pal8_320x240", True, 320, 240, 4, 14, 150, (40, 40, 20))
