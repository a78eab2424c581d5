"""Which frames of a C5 rank's corpus fail on the GPU, and what does the oracle say about them?  (debug aid)"""
import sys, os, argparse
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jsplayer_b200 import BatchDecoder, _lib
from oracle import pyoracle as O
rank = int(sys.argv[1]) if len(sys.argv) > 1 else 3
files = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wl = bench.C5(files)
specs = wl.specs(rank)
bd = BatchDecoder(insignificant_lines=36)
bd.configure(specs, pinned=True)
bd.upload(); bd.run(); bd.sync()
flags = bd.results()
first = np.cumsum([0] + [sp.n_frames for sp in specs])
bad = np.nonzero(flags & 4)[0]
print("errors:", len(bad))
for g in bad:
    s = int(np.searchsorted(first, g, side="right") - 1)
    sp = specs[s]
    fi, lo, hi = wl.where[s]
    fr = bench.spec_frames(sp)
    f = int(g - first[s])
    print("spec", s, "file", fi, "frames", lo, hi, "codec", int(sp.codec), "sp_version", sp.sp_version, "frame", f, "len", len(fr[f]), "head", fr[f][:2].hex(), "key", sp.keys[f])
for s in sorted({int(np.searchsorted(first, g, side="right") - 1) for g in bad}):
    sp = specs[s]
    fr = bench.spec_frames(sp)
    exp, ch, sg, st = O.decode_stream(int(sp.codec), sp.width, sp.height, sp.bpp, fr, keys=sp.keys, palette=sp.palette, insignificant_lines=36)
    print("spec", s, "oracle status", list(st))
    outs = [None] * bd.n_frames
    for f in range(sp.n_frames):
        outs[first[s] + f] = np.empty((sp.height, sp.width), dtype=np.int32)
    bd.download(outs)
    for f in range(sp.n_frames):
        eq = (outs[first[s] + f] == exp[f]).all()
        print("   frame", f, "gpu flags", int(flags[first[s] + f]), "equal to oracle:", bool(eq))
bd.close()
