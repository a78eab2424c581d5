"""Fuzz campaign for the ScreenPressor decoders: bit flips and truncations in I and P frames of synthetic streams (both
coders, 24 and 16 bpp, picture sizes down to 5 pixels wide); every frame of every corrupted stream must come out of the
CUDA path exactly as the CPU oracle defines it -- picture, error status, changed flag.  48 corrupted copies of a stream
form one batch.  Usage (GPU box): python tools/sp_fuzz.py [campaign 0..n]; prints the mismatching cases and their count.
Round 1 found with it: unreported late failures and JavaScript-double arithmetic after a failed symbol in the oracle,
a stale tile column and the out-of-picture predictor chain in the kernels."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
from oracle import pyoracle as O
SP = CodecType.codec_screenpressor
nbad = 0
SEED = int(sys.argv[1]) if len(sys.argv) > 1 else 0
SIZES = ((100, 60), (33, 17), (64, 48), (16, 16), (17, 33), (48, 16), (250, 130), (320, 240)) if SEED == 0 else ((9, 20), (12, 12), (5, 40), (31, 31), (15, 64), (96, 32), (640, 48), (200, 100))
for version, bpp in ((2, 24), (3, 24), (4, 24), (2, 16)):
    rng = np.random.default_rng(7000 + version + bpp + 1000 * SEED)
    for (w, h) in SIZES:
        frames, keys, pics = synth.sp_stream(w, h, 6, seed=w * 3 + version + 17 * SEED, version=version, gop=4, change_permille=200, bpp=bpp)
        specs, cases = [], []
        for trial in range(48):
            bad = []
            for fi, f in enumerate(frames):
                b = bytearray(f)
                mode = trial % 4
                hit = (fi > 0) if mode < 2 else (rng.random() < 0.5)
                if hit and len(b) > 6:
                    for _ in range(1 + trial % 5):
                        b[int(rng.integers(1, len(b)))] ^= int(1 << rng.integers(0, 8))
                    if mode == 3 and rng.random() < 0.3:
                        b = b[: int(rng.integers(2, len(b)))]
                bad.append(bytes(b))
            specs.append(StreamSpec(SP, w, h, bpp, frames=bad, keys=keys)); cases.append(bad)
        bd = BatchDecoder(insignificant_lines=16, significance=True); bd.configure(specs)
        outs, flags = bd.decode_host(); bd.close()
        k = 0
        for trial, bad in enumerate(cases):
            exp, ch, sg, st = O.decode_stream(O.CODEC_SCREENPRESSOR, w, h, bpp, bad, keys=keys, insignificant_lines=16)
            n = len(bad)
            gerr = [int(bool(f & _lib.JSP_FRAME_ERROR)) for f in flags[k:k + n]]
            same = [int((outs[k + i] == exp[i]).all()) for i in range(n)]
            gch = [int(bool(f & _lib.JSP_FRAME_CHANGED)) for f in flags[k:k + n]]
            okflags = all(gerr[i] or gch[i] == int(bool(ch[i])) for i in range(n))
            if gerr != [int(x != 0) for x in st] or not all(same) or not okflags:
                nbad += 1
                print(version, bpp, w, h, trial, "oracle", [int(x) for x in st], "gpu", gerr, "same", same, "changed", gch, [int(bool(x)) for x in ch], flush=True)
            k += n
        print("  ", version, bpp, w, h, "done", flush=True)
print("mismatches", nbad)
