"""A CPU model of the MSVideo1 kernel's entry tracking (jsplayer_b200/csrc/msv1_decode.cu, scan 2): the warp's 32 segment maps
are resolved in groups of 8 -- lane 8g + e follows entry e (every lane also entry 8) through the group's 8 segments and records
the entry INTO each segment as a nibble; group exits chain to the warp's map and, once the warp's entry is known, to the group
entries.  The model restates those steps lane by lane and must agree with the plain sequential walk on every warp entry, with
TERM (15) absorbing -- on random maps and on the maps of real synthetic frames.  No GPU."""
import numpy as np
import pytest

import synth

TERM = 15


def nib(m, e):
    return (m >> (4 * e)) & 15


def pack(map9):
    m = 0xF << 60                                          # MAP_TERM: nibble 15 maps to 15
    for e, v in enumerate(map9):
        m |= v << (4 * e)
    return m


def grouped_tracking(M, warp_entry):
    """-> (warp map as 9 exits, entry of each of the 32 lanes) the way the kernel computes them"""
    t_e, t_8, inter, inter8 = [0] * 32, [0] * 32, [0] * 32, [0] * 32
    for lane in range(32):
        grp, sub = lane & 24, lane & 7
        te, t8, rec, rec8 = sub, 8, 0, 0
        for i in range(8):
            Mi = M[grp + i]
            rec |= te << (4 * i); rec8 |= t8 << (4 * i)
            te, t8 = nib(Mi, te), nib(Mi, t8)
        t_e[lane], t_8[lane], inter[lane], inter8[lane] = te, t8, rec, rec8
    wmap = []
    for e in range(9):                                     # lanes 0..8 walk the four group exits
        cur = e
        for g in range(4):
            a, b = t_e[8 * g + (cur & 7)], t_8[8 * g]
            cur = TERM if cur == TERM else (a if cur < 8 else b)
        wmap.append(cur)
    entries = []
    for lane in range(32):
        eg, mine = warp_entry, warp_entry
        for g in range(4):
            if g == lane >> 3:
                mine = eg
            a, b = t_e[8 * g + (eg & 7)], t_8[8 * g]
            eg = TERM if eg == TERM else (a if eg < 8 else b)
        pk = inter[(lane & 24) + (mine & 7)]
        entries.append(TERM if mine == TERM else ((pk if mine < 8 else inter8[lane]) >> (4 * (lane & 7))) & 15)
    return wmap, entries


def sequential(M, warp_entry):
    e, entries = warp_entry, []
    for i in range(32):
        entries.append(e)
        e = TERM if e == TERM else nib(M[i], e)
    return e, entries


def check(maps):
    M = [pack(m) for m in maps]
    for we in list(range(9)) + [TERM]:
        wmap, entries = grouped_tracking(M, we)
        exit_seq, entries_seq = sequential(M, we)
        assert entries == entries_seq, (we, entries, entries_seq)
        if we != TERM:
            assert wmap[we] == exit_seq


def test_random_maps_with_terminators():
    rng = np.random.default_rng(11)
    for trial in range(60):
        p_term = [0.0, 0.02, 0.3][trial % 3]
        maps = [[TERM if rng.random() < p_term else int(rng.integers(0, 9)) for _ in range(9)] for _ in range(32)]
        check(maps)
    check([[8] * 9 for _ in range(32)])                    # every chain leaves through a 9-word opcode at word 15
    check([list(range(9)) for _ in range(32)])             # identity maps: the entry survives the whole warp


@pytest.mark.parametrize("mix", [(25, 50, 25), (0, 100, 0), (0, 0, 100)])
def test_maps_of_synthetic_frames(mix):
    f = synth.msv1_frame(False, 320, 240, 5, mix=mix)
    words = np.frombuffer(f[: len(f) // 2 * 2], dtype="<u2")
    G = np.concatenate([(words >> 15) & 1, np.zeros(32, dtype=words.dtype)])
    nseg = len(words) // 16

    def seg_map(s):
        m = []
        for e in range(9):
            p = s * 16 + e
            while p < s * 16 + 16:
                p += 1 if G[p] else (9 if G[p + 1] else 3)          # MSVideo1.hx:131-181
            m.append(p - (s * 16 + 16))
        return m
    maps = [seg_map(s) for s in range(nseg)]
    for w0 in range(0, min(nseg - 31, 32 * 12), 32):
        check(maps[w0:w0 + 32])
