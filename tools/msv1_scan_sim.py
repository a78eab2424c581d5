#!/usr/bin/env python
"""CPU simulation of the MSVideo1 opcode scan's entry hand-over (jsplayer_b200/csrc/msv1_decode.cu, scan 2) on the benchmark's
own synthetic frames: how often a 16-word segment's entry -> exit map is constant, and how many lane-to-lane shuffle rounds
the ROUND-2 kernel before the grouped entry tracking needed per warp and tile.  No GPU.

usage: PYTHONPATH=. python tools/msv1_scan_sim.py [mix, e.g. 25,50,25]

    mix       const maps   rounds before / after the look-back   32-step fallback
    25,50,25     14 %            12.3 / 6.1                            0.8 % of the warps
    0,0,100       6 %            12.8 / 13.1                           13 %
    0,100,0       0 %             1.0 / 31.0                          100 %
"""
import sys

import numpy as np

import synth

w, h = 1920, 1080
mix = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (25, 50, 25)
f = synth.msv1_frame(False, w, h, 7, mix=mix)
words = np.frombuffer(f[: len(f) // 2 * 2], dtype="<u2")
n = len(words)
G = np.concatenate([(words >> 15) & 1, np.zeros(32, dtype=words.dtype)])      # bit 15 of every word (zero past the end)
nseg = (n + 15) // 16


def seg_map(s):
    """exit offset (into the next segment) of the opcode chain that enters segment s at word e, e = 0..8"""
    m = []
    for e in range(9):
        p = s * 16 + e
        while p < s * 16 + 16:
            p += 1 if G[p] else (9 if G[p + 1] else 3)      # 1 colour / 8 colours / 2 colours (MSVideo1.hx:131-181)
        m.append(p - (s * 16 + 16))
    return m


maps = [seg_map(s) for s in range(nseg)]
const = [len(set(m)) == 1 for m in maps]
print("segments", nseg, "constant maps %.1f %%" % (100 * np.mean(const)))
r1, r2, fallback = [], [], 0
for w0 in range(0, nseg - 31, 32):
    M, C = maps[w0:w0 + 32], const[w0:w0 + 32]
    known = [False] + [C[i - 1] for i in range(1, 32)]
    entry = [0] + [M[i - 1][0] for i in range(1, 32)]
    r = 0
    while True:                                             # loop 1: a lane learns its entry from a predecessor whose exit is known
        r += 1
        exk = [known[i] or C[i] for i in range(32)]
        ex = [M[i][entry[i]] if known[i] else M[i][0] for i in range(32)]
        nk, ne, newly = known[:], entry[:], False
        for i in range(1, 32):
            if not known[i] and exk[i - 1]:
                nk[i], ne[i], newly = True, ex[i - 1], True
        known, entry = nk, ne
        if not newly:
            break
    r1.append(r)
    fallback += not (known[31] or C[31])
    known[0] = True                                         # loop 2: the warp's own entry has arrived through the look-back
    k = 0
    while not all(known):
        k += 1
        known = [known[i] or (i > 0 and known[i - 1]) for i in range(32)]
    r2.append(k)
print("warps", len(r1), "rounds before the look-back %.1f (max %d), after %.1f (max %d), 32-step fallback in %.1f %% of the warps"
      % (np.mean(r1), max(r1), np.mean(r2), max(r2), 100.0 * fallback / len(r1)))
