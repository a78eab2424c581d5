"""Fuzz campaign for the MSVideo1 decoder: bit flips, random words and truncations in key and inter frames of synthetic
streams (RGB555 and 8-bit, odd sizes, multi-tile frames); every frame must come out of the CUDA path exactly as the CPU
oracle has it (the block area of the picture, changed / error flags).  Usage (GPU box): python tools/msv1_fuzz.py [campaign]."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import BatchDecoder, StreamSpec, CodecType, _lib
import synth
from oracle import pyoracle as O
SEED = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nbad = 0
for is8 in (False, True):
    rng = np.random.default_rng(900 + SEED * 10 + is8)
    for (w, h) in ((64, 48), (320, 240), (100, 60), (4, 4), (8, 200), (1280, 720), (36, 20)):
        pal = synth.random_palette(w) if is8 else None
        frames = [synth.msv1_frame(is8, w, h, 11 + SEED)] + [synth.msv1_frame(is8, w, h, 20 + i + SEED, skip_permille=200, mean_skip=9) for i in range(3)] + [b""]
        keys = [1, 0, 0, 0, 0]
        specs, cases = [], []
        for trial in range(24 if w < 1000 else 6):
            bad = []
            for f in frames:
                b = bytearray(f)
                if len(b) > 2:
                    mode = trial % 4
                    for _ in range(1 + trial % 6):
                        at = int(rng.integers(0, len(b)))
                        if mode == 0: b[at] ^= 1 << int(rng.integers(0, 8))
                        elif mode == 1: b[at] = int(rng.integers(0, 256))
                        elif mode == 2:
                            a2 = at & ~1
                            b[a2:a2 + 2] = bytes([int(rng.integers(0, 256)), int(rng.choice([0x84, 0x87, 0x80, 0x00, 0xFF, 0x90]))])[: len(b) - a2]
                        else:
                            b = b[: int(rng.integers(0, len(b)))]; break
                bad.append(bytes(b))
            codec = CodecType.codec_msvc8 if is8 else CodecType.codec_msvc16
            specs.append(StreamSpec(codec, w, h, 8 if is8 else 16, frames=bad, keys=keys, palette=pal)); cases.append(bad)
        bd = BatchDecoder(insignificant_lines=8, significance=True); bd.configure(specs)
        outs, flags = bd.decode_host(); bd.close()
        k = 0
        bh, bw = h & ~3, w & ~3
        for trial, bad in enumerate(cases):
            exp, ch, sg, st = O.decode_stream(O.CODEC_MSVC8 if is8 else O.CODEC_MSVC16, w, h, 8 if is8 else 16, bad, keys=keys, palette=pal, insignificant_lines=8)
            n = len(bad)
            same = [int((outs[k + i][:bh, :bw] == np.asarray(exp[i])[:bh, :bw]).all()) for i in range(n)]
            gch = [int(bool(f & _lib.JSP_FRAME_CHANGED)) for f in flags[k:k + n]]
            gsg = [int(bool(f & _lib.JSP_FRAME_SIGNIFICANT)) for f in flags[k:k + n]]
            if not all(same) or gch != [int(bool(x)) for x in ch] or any(gsg[i] != int(bool(sg[i])) for i in range(1, n)):
                nbad += 1
                print(is8, w, h, trial, "same", same, "changed", gch, [int(bool(x)) for x in ch], "signif", gsg, [int(bool(x)) for x in sg], flush=True)
            k += n
        print("  ", is8, w, h, "done", flush=True)
print("mismatches", nbad)
