"""Symbol-level ScreenPressor ENCODER built on the second reading's model objects (oracle/sp_naive.py) -- TEST
INFRASTRUCTURE ONLY.

The C synthetic encoder (synth/) only emits what a sane encoder would: it never chooses a predictor that reads outside
the picture, never a zero-length run, never predictor 3 in an I frame.  The reference decoder nevertheless has a defined
(JavaScript) result for those streams, and the CUDA path must reproduce it.  This module turns an explicit list of
symbols -- the exact call sequence of ScreenPressor.DecompressI / DecompressP (ScreenPressor.hx:117-295, :302-484) --
into a valid v2 (range coder) or v3/v4 (rANS) frame, so tests can hand-craft such streams.  It shares NO code with
synth/*.c: the models are sp_naive's, the coders are written here from SURVEY.md Appendix C / D
(carry-propagating LZMA-style range encoder; two-pass reverse byte-wise rANS with 131072-symbol blocks).

A "frame script" is a list of tuples:
    ("clr", cxi, sym)  ("n", ptype, sym)  ("p", prev_ptype, sym)  ("x", sym)  ("bt", sym)  ("bn", sym)
    ("sxy", k, sym)    ("mx", sym)        ("my", sym)             ("bool", flag)
"""
import copy

from . import sp_naive as N


class _ProbeRC(N.RangeCoder):
    """A RangeCoder whose get_freq() returns a chosen value and whose decode() only records the interval: drives
    DecodeVal / DecodeValUni (search + model update, RangeCoder.hx:51-130) for a symbol WE choose."""

    def __init__(self):
        N.RangeCoder.__init__(self)
        self.value = 0

    def get_freq(self, total_freq):
        return self.value

    def decode(self, cumFreq, freq, total_freq):
        self.last = (cumFreq, freq, total_freq)


class RCFrameEncoder:
    """EntroCoderRC on the encoder side + the range encoder of SURVEY.md Appendix C."""

    def __init__(self):
        self.ec = N.EntroCoderRC()
        self.ec.rc = _ProbeRC()
        self.ec.preinit()

    def renewI(self):
        self.ec.renewI()

    def begin(self):
        self.low, self.range, self.cache, self.cache_size, self.out = 0, 0xFFFFFFFF, 0, 1, bytearray()

    def _shift_low(self):
        if self.low < 0xFF000000 or self.low >= 1 << 32:
            carry = self.low >> 32
            self.out.append((self.cache + carry) & 0xFF)
            for _ in range(self.cache_size - 1):
                self.out.append((0xFF + carry) & 0xFF)
            self.cache_size = 0
            self.cache = (self.low >> 24) & 0xFF
        self.cache_size += 1
        self.low = (self.low & 0x00FFFFFF) << 8

    def _put(self, cum, freq, tot):
        r = self.range // tot
        self.low += r * cum
        self.range = r * freq
        while self.range < 1 << 24:
            self.range = (self.range << 8) & 0xFFFFFFFF
            self._shift_low()

    def _val(self, table, maxc, step, sym):
        self.ec.rc.value = sum(table[i] for i in range(sym))
        got = self.ec.rc.DecodeVal(table, maxc, step)
        assert got == sym, (got, sym)
        self._put(*self.ec.rc.last)

    def sym(self, s):
        ec, kind = self.ec, s[0]
        if kind == "clr":
            _, cxi, c = s
            ec._touch_row(cxi)
            off = cxi * ec.CNTABSZ
            ec.rc.value = sum(ec.cntab[off + 17 + j] for j in range(c))
            got = ec.rc.DecodeValUni(ec.cntab, off, ec.SC_STEP)
            assert got == c
            self._put(*ec.rc.last)
        elif kind == "n":
            self._val(ec.ntab[s[1]], 256, ec.SC_NSTEP, s[2])
        elif kind == "p":
            self._val(ec.ptypetab[s[1]], 6, ec.SC_UNSTEP, s[2])
        elif kind == "x":
            self._val(ec.xxtab, 256, ec.SC_XXSTEP, s[1])
        elif kind == "bt":
            self._val(ec.bttab, 5, ec.SC_BTSTEP, s[1])
        elif kind == "bn":
            self._val(ec.ntab2, 256, ec.SC_BTNSTEP, s[1])
        elif kind == "sxy":
            self._val(ec.sxytab[s[1]], 16, ec.SC_SXYSTEP, s[2])
        elif kind == "mx":
            self._val(ec.mvtab[0], N.MSR_X * 2, ec.SC_MSTEP, s[1])
        elif kind == "my":
            self._val(ec.mvtab[1], N.MSR_Y * 2, ec.SC_MSTEP, s[1])
        else:
            raise ValueError("the range coder has no %r symbol" % (kind,))

    def finish(self):
        for _ in range(5):
            self._shift_low()
        return bytes(self.out)


class ANSFrameEncoder:
    """EntroCoderANS on the encoder side + the two-pass rANS encoder of SURVEY.md Appendix D."""

    def __init__(self, f0):
        self.ec = N.EntroCoderANS(f0)

    def renewI(self):
        self.ec.renewI()

    def begin(self):
        self.items = []             # (start, freq) or ("raw", byte)

    def _fixed(self, t, c):
        rcv = N.DecReceiver()
        t.decode(t.getCumFreq(c), rcv)
        assert rcv.c == c
        self.items.append((rcv.cumFreq, rcv.freq))

    def sym(self, s):
        ec, kind = self.ec, s[0]
        if kind == "clr":
            _, cxi, c = s
            dcx = ec.cntab[cxi]
            if dcx.u[0] < 4:                                     # raw byte, then the list kinds learn it
                dcx.update(c)
                self.items.append(("raw", c))
                return
            lo, hi = 0, 4095                                      # symbol is monotone in someFreq: bisect on copies
            while lo < hi:
                mid = (lo + hi) // 2
                probe = copy.deepcopy(dcx)
                probe.decode(mid)
                if N.Context.rcv.c < c:
                    lo = mid + 1
                else:
                    hi = mid
            N.Context.rcv = N.DecReceiver()
            dcx.decode(lo)
            rcv = N.Context.rcv
            assert rcv.c == c, "symbol %d is not codable in this context (got %d)" % (c, rcv.c)
            assert rcv.cumFreq + rcv.freq <= 4096
            self.items.append((rcv.cumFreq, rcv.freq))
        elif kind == "n":
            self._fixed(ec.ntab[s[1]], s[2])
        elif kind == "p":
            self._fixed(ec.ptypetab[s[1]], s[2])
        elif kind == "x":
            self._fixed(ec.xxtab, s[1])
        elif kind == "bt":
            self._fixed(ec.bttab, s[1])
        elif kind == "bn":
            self._fixed(ec.ntab2, s[1])
        elif kind == "sxy":
            self._fixed(ec.sxytab[s[1]], s[2])
        elif kind == "mx":
            self._fixed(ec.mvtab[0], s[1])
        elif kind == "my":
            self._fixed(ec.mvtab[1], s[1])
        elif kind == "bool":
            self.items.append((2048 if s[1] else 0, 2048))
        else:
            raise ValueError(kind)

    def finish(self):
        out = bytearray()
        B, L = N.Rans.B, N.Rans.RANS_BYTE_L
        blocks = [self.items[i:i + B] for i in range(0, len(self.items), B)] or [[]]
        if self.items and len(self.items) % B == 0:
            blocks.append([])                                    # the decoder re-reads a state after exactly B symbols
        for blk in blocks:
            rev = bytearray()
            x = L
            for it in reversed(blk):
                if it[0] == "raw":
                    rev.append(it[1])
                    continue
                start, freq = it
                x_max = ((L >> 12) << 8) * freq
                while x >= x_max:
                    rev.append(x & 0xFF)
                    x >>= 8
                x = ((x // freq) << 12) + (x % freq) + start
            rev += bytes([(x >> 24) & 0xFF, (x >> 16) & 0xFF, (x >> 8) & 0xFF, x & 0xFF])
            out += rev[::-1]
        return bytes(out)


class StreamBuilder:
    """Keeps one coder (models persist between frames like the decoder's) and wraps payloads into frames."""

    def __init__(self, version):
        self.version = version
        self.enc = RCFrameEncoder() if version == 2 else ANSFrameEncoder(64 if version == 3 else 32)

    def iframe(self, script):
        self.enc.renewI()
        self.enc.begin()
        for s in script:
            self.enc.sym(s)
        return bytes([((self.version - 1) << 4) | 2]) + self.enc.finish()

    def pframe(self, script):
        self.enc.begin()
        for s in script:
            self.enc.sym(s)
        return bytes([1]) + self.enc.finish()


class Scripter:
    """Builds frame scripts while tracking the colour-context state exactly as the decoder does
    (cx / cx1: ScreenPressor.hx:173-183, :274-275)."""

    def __init__(self, bpp=24, version=2):
        self.sh = 0 if (bpp == 16 and version == 2) else 2
        self.m1, self.s1, self.s = (0xFF00, 2, 16) if (bpp == 16 and version == 2) else (0xFC00, 4, 18)
        self.cx = self.cx1 = 0
        self.out = []

    def reset_ctx(self):
        self.cx = self.cx1 = 0

    def rgb(self, clr):
        for ch, v in enumerate((clr & 0xFF, (clr >> 8) & 0xFF, (clr >> 16) & 0xFF)):
            self.out.append(("clr", ch * 4096 + self.cx + self.cx1, v))
            self.cx1 = (self.cx << 6) & 0xFC0
            self.cx = v >> self.sh

    def after_run(self, clr):
        self.cx1 = (clr & self.m1) >> self.s1
        self.cx = clr >> self.s

    def take(self):
        o, self.out = self.out, []
        return o
