"""Hand-crafted ScreenPressor streams whose syntax no sane encoder emits but the reference decoder has a deterministic
(JavaScript) result for.  Built with the symbol-level encoder of oracle/sp_naive_enc.py (second reading's models; shares
no code with synth/*.c).  Used by tests/test_sp_second_reading.py on the CPU (oracle vs second reading) and under -m gpu
(CUDA path vs second reading)."""
import numpy as np

from oracle.sp_naive_enc import StreamBuilder, Scripter


def iframe_script(sc, px, X, extra=False):
    """Whole picture with predictor-0 runs only.  extra=True sprinkles legal no-ops between the runs: zero-length runs
    and predictor 3, which has no case in DecompressI's switch (ScreenPressor.hx:242-273) and writes nothing."""
    flat = [int(v) for v in px.reshape(-1)]
    end = len(flat)
    di = 0
    sc.reset_ctx()
    while di < X + 1:
        c = flat[di]
        n = 1
        while n < 255 and di + n < end and flat[di + n] == c:
            n += 1
        sc.rgb(c)
        sc.out.append(("n", 0, n))
        di += n
    ptype = 0
    k = 0
    while di < end:
        c = flat[di]
        n = 1
        while n < 255 and di + n < end and flat[di + n] == c:
            n += 1
        if extra and k % 5 == 1:                     # predictor 3: n is decoded, nothing written, context from old clr
            sc.out.append(("p", ptype, 3))
            sc.out.append(("n", 3, 7))
            ptype = 3
            sc.after_run(flat[di - 1])
        if extra and k % 7 == 2:                     # zero-length run of predictor 2: clr unchanged
            sc.out.append(("p", ptype, 2))
            sc.out.append(("n", 2, 0))
            ptype = 2
            sc.after_run(flat[di - 1])
        sc.out.append(("p", ptype, 0))
        ptype = 0
        sc.rgb(c)
        sc.out.append(("n", 0, n))
        di += n
        sc.after_run(c)
        k += 1
    return sc.take()


def _picture(X, Y, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 4, size=(Y, X)).astype(np.int32) * 0x203040 + 0x010203


def oob_predictors(version):
    """P frame whose data blocks use predictors 4, 5, 2, 1 on the picture's first rows and a block's first column:
    neighbours at negative indices are `undefined` (ScreenPressor.hx:440-449)."""
    X, Y = 40, 36
    px = _picture(X, Y, 5)
    sb, sc = StreamBuilder(version), Scripter(24, version)
    f0 = sb.iframe(iframe_script(sc, px, X))
    sc.reset_ctx()
    sc.out = [("x", 0), ("x", 0), ("x", 3), ("x", 0), ("bt", 1), ("bn", 1), ("bt", 0), ("bn", 2), ("bt", 1), ("bn", 1)]

    def run(prev_pt, pt, n):
        sc.out.append(("p", prev_pt, pt))
        sc.out.append(("n", pt, n))
    run(0, 4, 16)                 # row 0: every neighbour outside
    run(4, 4, 16)                 # row 1: x = 0 has its above-left neighbour at index -1
    run(4, 5, 16)
    run(5, 2, 16)
    run(2, 1, 16)
    run(1, 3, 16 * 11)
    run(0, 1, 16)                 # block (0,1): predictor 1 at x = 0 reads the previous row's last pixel
    run(1, 4, 32)
    run(4, 5, 16)
    run(5, 3, 16 * 12)
    f1 = sb.pframe(sc.take())
    return X, Y, 24, [f0, f1], [1, 0]


def noop_runs(version):
    """I frame with zero-length runs and predictor-3 no-ops; P frame with zero-length runs inside a data block."""
    X, Y = 37, 21
    px = _picture(X, Y, 6)
    sb, sc = StreamBuilder(version), Scripter(24, version)
    f0 = sb.iframe(iframe_script(sc, px, X, extra=True))
    sc.reset_ctx()
    sc.out = [("x", 1), ("x", 0), ("x", 1), ("x", 0), ("bt", 1), ("bn", 1)]
    sc.out += [("p", 0, 3), ("n", 3, 0), ("p", 3, 2), ("n", 2, 0), ("p", 2, 0)]
    sc.rgb(0x556677)
    sc.out += [("n", 0, 200), ("p", 0, 1), ("n", 1, 0), ("p", 1, 1), ("n", 1, 56)]
    f1 = sb.pframe(sc.take())
    return X, Y, 24, [f0, f1], [1, 0]


def wild_motion_and_subrects(version):
    """Motion vectors whose source rows lie above the picture (negative indices -> 0) or wrap into the previous row, and a
    sub-rectangle in the last block column that runs past the right edge (its writes wrap into the next row)."""
    X, Y = 40, 36
    px = _picture(X, Y, 8)
    sb, sc = StreamBuilder(version), Scripter(24, version)
    f0 = sb.iframe(iframe_script(sc, px, X))
    sc.reset_ctx()
    #            xx1 = 0          xx2 = 5        blocks 0..5: motion, motion+sub, data+sub (last column), none, motion, none
    sc.out = [("x", 0), ("x", 0), ("x", 5), ("x", 0),
              ("bt", 3), ("bn", 1), ("bt", 4), ("bn", 1), ("bt", 2), ("bn", 1), ("bt", 0), ("bn", 1), ("bt", 3), ("bn", 1),
              ("bt", 0), ("bn", 1)]
    can_bool = version != 2

    def mv(mx, my, same=False):
        if can_bool:
            sc.out.append(("bool", 1 if same else 0))
            if same:
                return
        sc.out.append(("mx", mx + 256))
        sc.out.append(("my", my + 256))
    mv(-5, -3)                                                    # block 0: rows -3..12, columns -5..10
    sc.out += [("sxy", 0, 2), ("sxy", 1, 1), ("sxy", 2, 9), ("sxy", 3, 12)]
    mv(3, 20)                                                     # block 1 sub-rectangle, source partly below the picture
    sc.out += [("sxy", 0, 1), ("sxy", 1, 0), ("sxy", 2, 15), ("sxy", 3, 3)]   # block 2 (x16 = 32, X = 40): x 33..47, y 0..3
    sc.out += [("p", 0, 0)]
    sc.rgb(0x112233)
    sc.out += [("n", 0, 15 * 4)]
    if can_bool:
        mv(3, 20, same=True)                                      # block 4: "same vector as last time"
    else:
        mv(3, 20)
    f1 = sb.pframe(sc.take())
    return X, Y, 24, [f0, f1], [1, 0]


CRAFTED = {"oob_predictors": oob_predictors, "noop_runs": noop_runs, "wild_motion_and_subrects": wild_motion_and_subrects}
