"""SECOND, INDEPENDENT reading of the ScreenPressor decoder -- TEST INFRASTRUCTURE ONLY (never on the product path).

A deliberately naive, line-by-line Python restatement of the reference's

    /root/reference/src/ANS.hx          (whole file: Rans :5-49, FixedSizeRansCtx :54-145, SymbList/Cx1-3 :155-208,
                                         SmallContext :210-310, Cx4 :312-327, Cx5 :329-392, Cx6 :394-704, Cx7 :706-772,
                                         Context :785-860, Sorter :862-872)
    /root/reference/src/RangeCoder.hx   (whole file)
    /root/reference/src/EntroCoders.hx  (EntroCoderRC :31-180, EntroCoderANS :182-313)
    /root/reference/src/ScreenPressor.hx (ctor :53-66, initEntro :68-82, Preinit :89-92, IsKeyFrame :99-104,
                                         RenewI :111-115, DecompressI :117-295, DecompressP :302-484)

written WITHOUT looking at oracle/*.c, synth/ans_models.c or the CUDA kernels: one Python object per reference object,
the reference's process-global statics kept as class attributes, JavaScript typed-array semantics made explicit
(Uint8/Uint16 stores wrap, Int32Array stores of `undefined` give 0, out-of-range typed-array reads give `undefined`,
arithmetic on `undefined` gives NaN and NaN coerces to 0 under the bit operators).  It exists to pin the C oracle (and
through it the CUDA path): tests/golden/make_sp_naive_golden.py runs it over the synthetic corpus and commits pictures
plus per-symbol (c, freq, cumFreq) traces, tests/test_sp_second_reading.py asserts oracle == this == GPU.

It is slow (pure Python, ~1 us per byte-code) and only ever sees small frames.
"""

UNDEF = None          # JavaScript `undefined`

import collections
EVENTS = collections.Counter()      # instrumentation only (which rare model paths a corpus exercised); not in the reference


class JsSemanticsError(Exception):
    """Raised where the reference would compute with NaN / Infinity inside the coder state (corrupt streams only)."""


def _i32(x):
    x &= 0xFFFFFFFF
    return x - 0x100000000 if x & 0x80000000 else x


class U8:
    """js.lib.Uint8Array: zero initialised, stores wrap mod 256, out-of-range reads are `undefined`, writes dropped."""
    __slots__ = ("a",)

    def __init__(self, n_or_list):
        self.a = [0] * n_or_list if isinstance(n_or_list, int) else [v & 0xFF for v in n_or_list]

    @property
    def length(self):
        return len(self.a)

    def __getitem__(self, i):
        return self.a[i] if 0 <= i < len(self.a) else UNDEF

    def __setitem__(self, i, v):
        if 0 <= i < len(self.a):
            self.a[i] = v & 0xFF


class U16(U8):
    __slots__ = ()

    def __init__(self, n):
        self.a = [0] * n

    def __setitem__(self, i, v):
        if 0 <= i < len(self.a):
            self.a[i] = v & 0xFFFF


class U32(U8):
    __slots__ = ()

    def __init__(self, n):
        self.a = [0] * n

    def __setitem__(self, i, v):
        if 0 <= i < len(self.a):
            self.a[i] = v & 0xFFFFFFFF


def _num(v):
    """A typed-array element used in `|`, `<<`: undefined -> 0."""
    return 0 if v is UNDEF else v


# ------------------------------------------------------------------------------------------------ ANS.hx:5-49
class Rans:
    B = 131072
    PROB_SCALE = 4096
    RANS_BYTE_L = 1 << 23

    def __init__(self, srcdata, pos0=0):
        self.oob_reads = 0
        self.reinitImpl(srcdata, pos0)

    def reinit(self):
        EVENTS["rans_reinit"] += 1
        self.reinitImpl(self.data, self.pos)

    def _rd(self, i):
        v = self.data[i]
        if v is UNDEF:
            self.oob_reads += 1
            return 0
        return v

    def reinitImpl(self, srcdata, i):
        self.data = srcdata
        x = self._rd(i + 0)
        x |= self._rd(i + 1) << 8
        x |= self._rd(i + 2) << 16
        x = _i32(x | _i32(self._rd(i + 3) << 24))
        self.r = x
        self.pos = i + 4

    def decGet(self):
        return self.r & 4095

    def decAdvance(self, start, freq):
        x = self.r
        x = freq * (_i32(x) >> 12) + (x & 4095) - start
        while x < Rans.RANS_BYTE_L:
            x = _i32(_i32(x << 8) | self._rd(self.pos))
            self.pos += 1
        self.r = x

    def raw(self):
        v = self._rd(self.pos)
        self.pos += 1
        return v


class DecReceiver:
    __slots__ = ("c", "freq", "cumFreq")

    def __init__(self):
        self.c = 0
        self.freq = 0
        self.cumFreq = 0


# ------------------------------------------------------------------------------------------------ ANS.hx:54-145
class FixedSizeRansCtx:
    STEP_FX = 16
    step = STEP_FX
    Dshift = 7
    D = 1 << Dshift

    def __init__(self, NSymb):
        self.NSym = NSymb
        self.freqs = U16(NSymb * 2)
        self.cnts = U16(NSymb)
        self.decTable = U8(32)
        self.cntsum = 0

    def setFreq(self, i, fr, cf):
        self.freqs[i * 2] = fr
        self.freqs[i * 2 + 1] = cf

    def readFreq(self, i):
        return self.freqs[i * 2]

    def readCumFreq(self, i):
        return self.freqs[i * 2 + 1]

    getCumFreq = readCumFreq

    def incrCnt(self, c):
        step = FixedSizeRansCtx.step
        self.cnts[c] = self.cnts[c] + step
        self.cntsum += step
        if self.cntsum + step > Rans.PROB_SCALE:
            EVENTS["fixed_rebuild_%d" % self.NSym] += 1
            self.cntsum = 0
            cf = 0
            for j in range(self.NSym):
                fr = self.cnts[j]
                self.setFreq(j, fr, cf)
                k0 = (cf + FixedSizeRansCtx.D - 1) >> FixedSizeRansCtx.Dshift
                k1 = ((cf + fr - 1) >> FixedSizeRansCtx.Dshift) + 1
                for k in range(k0, k1):
                    self.decTable[k] = j
                cf += fr
                self.cnts[j] = self.cnts[j] - (fr >> 1)
                self.cntsum += self.cnts[j]

    def decode(self, someFreq, rcv):
        c0 = self.decTable[someFreq >> FixedSizeRansCtx.Dshift]
        for j in range(c0, self.NSym - 1):
            if self.getCumFreq(j + 1) > someFreq:
                rcv.freq = self.readFreq(j)
                rcv.cumFreq = self.readCumFreq(j)
                rcv.c = j
                self.incrCnt(j)
                return True
        rcv.freq = self.readFreq(self.NSym - 1)
        rcv.cumFreq = self.readCumFreq(self.NSym - 1)
        rcv.c = self.NSym - 1
        self.incrCnt(self.NSym - 1)
        return True

    def renew(self):
        cf = 0
        fr = Rans.PROB_SCALE // self.NSym
        c0 = fr - (fr >> 1)
        self.cntsum = c0 * self.NSym
        for i in range(self.NSym):
            self.setFreq(i, fr, cf)
            self.cnts[i] = c0
            k0 = (cf + FixedSizeRansCtx.D - 1) >> FixedSizeRansCtx.Dshift
            k1 = ((cf + fr - 1) >> FixedSizeRansCtx.Dshift) + 1
            for k in range(k0, k1):
                self.decTable[k] = i
            cf += fr


FOUND, ADDED, NOROOM = 0, 1, 2


# ------------------------------------------------------------------------------------------------ ANS.hx:155-208
class SymbList:
    def __init__(self, num):
        self.symb = U8(num)
        self.d = 0

    def findOrAdd(self, c):
        for i in range(self.d):
            if self.symb[i] == c:
                return FOUND
        if self.d < self.symb.length:
            self.symb[self.d] = c
            self.d += 1
            return ADDED
        return NOROOM


class Cx1(SymbList):
    def __init__(self, c):
        SymbList.__init__(self, 14)
        self.d = 1
        self.symb[0] = c


class Cx2(SymbList):
    def __init__(self, c1, c):
        SymbList.__init__(self, 64)
        for i in range(c1.d):
            self.symb[i] = c1.symb[i]
        self.symb[c1.d] = c
        self.d = c1.d + 1


class Cx3(SymbList):
    def __init__(self, c2, c):
        SymbList.__init__(self, 256)
        for i in range(c2.d):
            self.symb[i] = c2.symb[i]
        self.symb[c2.d] = c
        self.d = c2.d + 1


def _insort_view(u8, n):
    """Sorter.insort applied to `u8.subarray(0, n)`: a typed-array VIEW, so the parent array is sorted in place
    (ANS.hx:862-872 with :229-231 and :524-526)."""
    a = u8.a
    for i in range(1, n):
        j = i
        while j > 0 and a[j - 1] > a[j]:
            a[j], a[j - 1] = a[j - 1], a[j]
            j -= 1


# ------------------------------------------------------------------------------------------------ ANS.hx:210-310
class SmallContext:
    f0 = 50
    totFr = 0           # static var totFr (process-global in the reference)

    def __init__(self, size):
        self.S = size
        self.symbols = U8(size)
        self.freqs = U16(size)
        self.maxpos = 0
        self.d = 0

    def create(self, c1, c):
        self.d = c1.d
        _insort_view(c1.symb, self.d)
        ss = c1.symb
        for i in range(self.d):
            self.symbols[i] = ss[i]
            if self.symbols[i] == c:
                self.freqs[i] = 2 * SmallContext.f0
                self.maxpos = i
            else:
                self.freqs[i] = SmallContext.f0

    def addSymb(self, pos, c):
        if self.d == self.S:
            return False
        i = self.d - 1
        while i >= pos:
            self.symbols[i + 1] = self.symbols[i]
            self.freqs[i + 1] = self.freqs[i]
            i -= 1
        self.symbols[pos] = c
        self.freqs[pos] = SmallContext.f0
        self.d += 1
        if self.maxpos >= pos:
            self.maxpos += 1
        SmallContext.totFr += SmallContext.f0
        if SmallContext.totFr + SmallContext.f0 > Rans.PROB_SCALE:
            self.rescale()
        return True

    def rescale(self):
        EVENTS["small_rescale_S%d" % self.S] += 1
        s = 256 - self.d
        for i in range(self.d):
            self.freqs[i] = self.freqs[i] - (self.freqs[i] >> 1)
            s += self.freqs[i]
        SmallContext.totFr = s

    def decodeSC(self, someFreq, rcv, totFr0):
        f0 = SmallContext.f0
        SmallContext.totFr = totFr0
        shift = 0
        tot = totFr0
        while tot <= Rans.PROB_SCALE / 2:
            tot <<= 1
            shift += 1
        someFreq >>= shift
        bonus = (Rans.PROB_SCALE - tot) >> shift
        maxFreq = self.freqs[self.maxpos]
        self.freqs[self.maxpos] = self.freqs[self.maxpos] + bonus
        cumFr = 0
        lastSymb = 0
        pos = 0
        while pos < self.d:
            s = self.symbols[pos]
            startFr = cumFr + s - lastSymb
            if someFreq < startFr:
                rcv.c = someFreq - cumFr + lastSymb
                cumFr = someFreq
                rcv.cumFreq = cumFr << shift
                rcv.freq = 1 << shift
                self.freqs[self.maxpos] = maxFreq
                return self.addSymb(pos, rcv.c)
            fr = self.freqs[pos]
            if startFr + fr > someFreq:
                rcv.c = s
                cumFr += rcv.c - lastSymb
                rcv.cumFreq = cumFr << shift
                rcv.freq = fr << shift
                self.freqs[self.maxpos] = maxFreq
                self.freqs[pos] = self.freqs[pos] + f0
                SmallContext.totFr += f0
                if pos != self.maxpos and self.freqs[pos] > self.freqs[self.maxpos]:
                    self.maxpos = pos
                if SmallContext.totFr + f0 > Rans.PROB_SCALE:
                    self.rescale()
                return True
            cumFr += s - lastSymb + fr
            lastSymb = s + 1
            pos += 1
        self.freqs[self.maxpos] = maxFreq
        if pos == self.d:
            rcv.c = lastSymb + someFreq - cumFr
            rcv.cumFreq = someFreq << shift
            rcv.freq = 1 << shift
            return self.addSymb(pos, rcv.c)
        return True


# ------------------------------------------------------------------------------------------------ ANS.hx:312-327
class Cx4(SmallContext):
    def __init__(self, c1, c):
        SmallContext.__init__(self, 4)
        self.create(c1, c)

    def decode(self, someFreq, rcv):
        f = self.freqs
        totFr = f[0] + f[1] + f[2] + f[3] + 256 - self.d
        return self.decodeSC(someFreq, rcv, totFr)

    def upgrade(self, c):
        return (5, Cx5.fromCx4(self, c))


# ------------------------------------------------------------------------------------------------ ANS.hx:329-392
class Cx5(SmallContext):
    def __init__(self):
        SmallContext.__init__(self, 16)
        self.cntsum = 0

    @staticmethod
    def fromCx1(c1, c):
        cx = Cx5()
        cx.create(c1, c)
        cx.calcSum()
        return cx

    @staticmethod
    def fromCx4(c4, c):
        cx = Cx5()
        cx.createFrom4(c4, c)
        return cx

    def createFrom4(self, c4, c):
        i = 0
        dd = c4.d
        totFr = 0
        while i < dd and c4.symbols[i] < c:
            self.symbols[i] = c4.symbols[i]
            self.freqs[i] = c4.freqs[i]
            totFr += self.freqs[i]
            i += 1
        j = i
        self.symbols[j] = c
        self.freqs[j] = SmallContext.f0
        totFr += self.freqs[j]
        j += 1
        while i < dd:
            self.symbols[j] = c4.symbols[i]
            self.freqs[j] = c4.freqs[i]
            totFr += self.freqs[j]
            i += 1
            j += 1
        self.d = dd + 1
        if totFr > Rans.PROB_SCALE:
            self.rescale()
        self.calcSum()

    def calcSum(self):
        totFr = 256 - self.d
        for i in range(self.d):
            totFr += self.freqs[i]
        self.cntsum = totFr

    def decode(self, someFreq, rcv):
        res = self.decodeSC(someFreq, rcv, self.cntsum)
        self.cntsum = SmallContext.totFr
        return res

    def upgrade(self, c):
        cx = Cx6()
        cx.createFrom5(self, c)
        return (6, cx)


# ------------------------------------------------------------------------------------------------ ANS.hx:394-704
class Cx6:
    _cnts = U16(256)        # static scratch (ANS.hx:401-402)
    _freqs = U16(512)
    Step = 25
    f0 = 32                 # static; set by EntroCoderANS's constructor (EntroCoders.hx:210)

    def __init__(self):
        self.symbols = None
        self.freqs = None
        self.cnts = None
        self.d = 0
        self.fshift = 0

    def setFreq(self, i, fr, cf):
        self.freqs[i * 2] = fr
        self.freqs[i * 2 + 1] = cf

    def readFreq(self, idx):
        return self.freqs[idx * 2]

    def readCumFreq(self, idx):
        return self.freqs[idx * 2 + 1]

    def init(self, S):
        self.symbols = U8(S)
        self.freqs = U16(S * 2)
        self.cnts = U16(S + 1)

    def createFrom5(self, c5, c):
        self.init(32)
        S = 32
        oldd = c5.d
        totFr = 256 - oldd
        for i in range(oldd):
            totFr += c5.freqs[i]
        shift = 0
        tot = totFr
        while tot <= Rans.PROB_SCALE / 2:
            tot <<= 1
            shift += 1
        cumFr = 0
        lastSymb = 0
        for pos in range(oldd):
            s = c5.symbols[pos]
            cumFr += s - lastSymb
            cfr = c5.freqs[pos]
            fr = cfr << shift
            self.setFreq(pos, fr, cumFr << shift)
            self.cnts[pos] = fr - (fr >> 1)
            self.symbols[pos] = s
            cumFr += cfr
            lastSymb = s + 1
        self.fshift = shift
        fr_freq = 1 << self.fshift
        fr_cumFreq = 0
        if c > 0:
            lowerSym = -1
            lfreq = 0
            lcumFreq = 0
            for i in range(oldd):
                s = self.symbols[i]
                if s > lowerSym and s < c:
                    lowerSym = s
                    lfreq = self.readFreq(i)
                    lcumFreq = self.readCumFreq(i)
            if lfreq > 0:
                fr_cumFreq = lcumFreq + lfreq + ((c - lowerSym - 1) << self.fshift)
            else:
                fr_cumFreq = c << self.fshift
        self.setFreq(oldd, fr_freq, fr_cumFreq)
        self.cnts[oldd] = fr_freq - (fr_freq >> 1)
        self.symbols[oldd] = c
        self.d = oldd + 1
        step = Cx6.Step << self.fshift
        self.cnts[oldd] = self.cnts[oldd] + step
        self.cnts[S] = self.cnts[S] + step
        if self.cnts[S] + step > Rans.PROB_SCALE:
            self.rescaleDec()
        self.calcSum()
        for i in range(0, self.d - 1):
            for j in range(i + 1, self.d):
                fj = self.readFreq(j)
                fi = self.readFreq(i)
                if fj > fi:
                    cfi = self.readCumFreq(i)
                    cfj = self.readCumFreq(j)
                    self.setFreq(i, fj, cfj)
                    self.setFreq(j, fi, cfi)
                    tc = self.cnts[i]
                    self.cnts[i] = self.cnts[j]
                    self.cnts[j] = tc
                    ts = self.symbols[i]
                    self.symbols[i] = self.symbols[j]
                    self.symbols[j] = ts

    def createFrom2(self, cx, c):
        S0 = 32 if cx.d <= 32 else 64
        self.init(S0)
        f0 = Cx6.f0
        oldd = cx.d
        totFr = 256 - oldd
        totFr += oldd * f0 + f0
        shift = 0
        tot = totFr
        while tot <= Rans.PROB_SCALE / 2:
            tot <<= 1
            shift += 1
        cumFr = 0
        lastSymb = 0
        _insort_view(cx.symb, oldd)
        newSymbPos = 0
        for pos in range(oldd):
            s = cx.symb[pos]
            cumFr += s - lastSymb
            if s == c:
                newSymbPos = pos
                cfr = f0 * 2
            else:
                cfr = f0
            fr = cfr << shift
            self.setFreq(pos, fr, cumFr << shift)
            self.symbols[pos] = s
            self.cnts[pos] = fr - (fr >> 1)
            cumFr += cfr
            lastSymb = s + 1
        self.d = oldd
        self.fshift = shift
        self.calcSum()
        if newSymbPos > 0:
            fr0 = self.readFreq(0)
            cf0 = self.readCumFreq(0)
            frc = self.readFreq(newSymbPos)
            cfc = self.readCumFreq(newSymbPos)
            self.setFreq(0, frc, cfc)
            self.setFreq(newSymbPos, fr0, cf0)
            sym0 = self.symbols[0]
            cnt0 = self.cnts[0]
            cntc = self.cnts[newSymbPos]
            self.cnts[0] = cntc
            self.cnts[newSymbPos] = cnt0
            self.symbols[0] = c
            self.symbols[newSymbPos] = sym0

    def calcSum(self):
        shft = self.fshift - 1 if self.fshift > 0 else 0
        sum_ = (256 - self.d) << shft
        S = self.symbols.length
        for i in range(S):
            sum_ += self.cnts[i]
        self.cnts[S] = sum_

    def rescaleDec(self):
        EVENTS["cx6_rescaleDec"] += 1
        _cnts = Cx6._cnts
        _freqs = Cx6._freqs
        sh = self.fshift - 1 if self.fshift > 0 else 0
        c0 = 1 << sh
        for i in range(256):
            _cnts[i] = c0
        for i in range(self.d):
            _cnts[self.symbols[i]] = self.cnts[i]
        cumFr = 0
        for i in range(256):
            _freqs[i * 2] = _cnts[i]
            _freqs[i * 2 + 1] = cumFr
            cumFr += _cnts[i]
        if self.fshift > 0:
            self.fshift -= 1
        shft = self.fshift - 1 if self.fshift > 0 else 0
        cntsum = (256 - self.d) << shft
        for i in range(self.d):
            self.cnts[i] = self.cnts[i] - (self.cnts[i] >> 1)
            cntsum += self.cnts[i]
            idx = self.symbols[i]
            self.setFreq(i, _freqs[idx * 2], _freqs[idx * 2 + 1])
        self.cnts[self.symbols.length] = cntsum

    def decode(self, someFreq, rcv):
        lfreq = 0
        lcumFreq = 0
        lowerSym = 0
        for i in range(self.d):
            cf = self.readCumFreq(i)
            if cf <= someFreq:
                fr = self.readFreq(i)
                if cf + fr > someFreq:
                    rcv.c = self.symbols[i]
                    rcv.freq = fr
                    rcv.cumFreq = cf
                    self.incrCntDec(i)
                    return True
                if cf >= lcumFreq:
                    lfreq = fr
                    lcumFreq = cf
                    lowerSym = self.symbols[i]
        fr_freq = 1 << self.fshift
        if lfreq > 0:
            cumFr = lcumFreq + lfreq
            x = (someFreq - cumFr) >> self.fshift
            c = x + lowerSym + 1
            fr_cumFreq = lcumFreq + lfreq + (x << self.fshift)
        else:
            c = someFreq >> self.fshift
            fr_cumFreq = c << self.fshift
        rcv.freq = fr_freq
        rcv.cumFreq = fr_cumFreq
        rcv.c = c
        p = self.addDec(c, fr_freq, fr_cumFreq)
        if p < 0:
            if self.symbols.length == 64:
                return False
            self.growDec()
            p = self.addDec(c, fr_freq, fr_cumFreq)
        self.incrCntDec(p)
        return True

    def addDec(self, c, freq, cumFreq):
        if self.d >= 40 or self.d >= self.symbols.length:
            return -1
        pos = self.d
        self.symbols[pos] = c
        self.setFreq(pos, freq, cumFreq)
        self.cnts[pos] = freq - (freq >> 1)
        self.d += 1
        return pos

    def growDec(self):
        EVENTS["cx6_grow"] += 1
        S = self.symbols.length * 2
        sym = U8(S)
        cs = U16(S + 1)
        fs = U16(S * 2)
        for i in range(self.d):
            sym[i] = self.symbols[i]
            cs[i] = self.cnts[i]
            fs[i * 2] = self.freqs[i * 2]
            fs[i * 2 + 1] = self.freqs[i * 2 + 1]
        cs[S] = self.cnts[self.symbols.length]
        self.symbols = sym
        self.cnts = cs
        self.freqs = fs

    def incrCntDec(self, pos):
        step = Cx6.Step << self.fshift
        S = self.symbols.length
        self.cnts[pos] = self.cnts[pos] + step
        self.cnts[S] = self.cnts[S] + step
        if pos > 0 and self.cnts[pos] > self.cnts[pos - 1]:
            tc = self.cnts[pos]
            self.cnts[pos] = self.cnts[pos - 1]
            self.cnts[pos - 1] = tc
            fp = self.readFreq(pos)
            cfp = self.readCumFreq(pos)
            self.setFreq(pos, self.readFreq(pos - 1), self.readCumFreq(pos - 1))
            self.setFreq(pos - 1, fp, cfp)
            ts = self.symbols[pos]
            self.symbols[pos] = self.symbols[pos - 1]
            self.symbols[pos - 1] = ts
        if self.cnts[S] + step > Rans.PROB_SCALE:
            self.rescaleDec()

    def upgrade(self, c):
        cx = Cx7()
        cx.createFrom6(self, c)
        return (7, cx)


# ------------------------------------------------------------------------------------------------ ANS.hx:706-772
class Cx7(FixedSizeRansCtx):
    def __init__(self):
        FixedSizeRansCtx.__init__(self, 256)

    def createFrom3(self, c3, c):
        for i in range(256):
            self.freqs[i * 2] = 1
            self.cnts[i] = 1
        d = c3.d
        f0 = (Rans.PROB_SCALE - (256 - d)) // (d + 1)
        c0 = f0 - (f0 >> 1)
        for i in range(d):
            s = c3.symb[i]
            self.freqs[s * 2] = f0
            self.cnts[s] = c0
        self.freqs[c * 2] = self.freqs[c * 2] + f0
        self.cnts[c] = self.cnts[c] + FixedSizeRansCtx.step
        self.cntsum = 0
        cf = 0
        for i in range(256):
            self.cntsum += self.cnts[i]
            self.freqs[i * 2 + 1] = cf
            fr = self.freqs[i * 2]
            k0 = (cf + FixedSizeRansCtx.D - 1) >> FixedSizeRansCtx.Dshift
            k1 = ((cf + fr - 1) >> FixedSizeRansCtx.Dshift) + 1
            for k in range(k0, k1):
                self.decTable[k] = i
            cf += fr

    def createFrom6(self, c6, c):
        S = c6.symbols.length
        self.cntsum = c6.cnts[S]
        for i in range(S):
            if c6.cnts[i] > 0:
                x = c6.symbols[i]
                self.setFreq(x, c6.freqs[i * 2], c6.freqs[i * 2 + 1])
                self.cnts[x] = c6.cnts[i]
        funmet = 1 << c6.fshift
        cntUnmet = funmet - (funmet >> 1)
        cumFr = 0
        for i in range(256):
            if self.freqs[i * 2] > 0:
                fr = self.freqs[i * 2]
            else:
                self.setFreq(i, funmet, cumFr)
                self.cnts[i] = cntUnmet
                fr = funmet
            k0 = (cumFr + FixedSizeRansCtx.D - 1) >> FixedSizeRansCtx.Dshift
            k1 = ((cumFr + fr - 1) >> FixedSizeRansCtx.Dshift) + 1
            for k in range(k0, k1):
                self.decTable[k] = i
            cumFr += fr


# ------------------------------------------------------------------------------------------------ ANS.hx:774-860
class Context:
    rcv = None              # public static var rcv

    def __init__(self):
        self.u = (0, None)              # KindNone
        Context.rcv = DecReceiver()     # yes: every constructor replaces the static receiver (ANS.hx:789)

    def renew(self):
        self.u = (0, None)

    def decode(self, someFreq):
        kind, x = self.u
        rcv = Context.rcv
        if kind == 6:
            if not x.decode(someFreq, rcv):
                self.u = x.upgrade(rcv.c)
        elif kind == 7:
            x.decode(someFreq, rcv)
        elif kind == 4:
            if not x.decode(someFreq, rcv):
                self.u = x.upgrade(rcv.c)
        elif kind == 5:
            if not x.decode(someFreq, rcv):
                self.u = x.upgrade(rcv.c)
        else:
            return False
        return True

    def update(self, c):
        kind, x = self.u
        if kind == 0:
            self.u = (1, Cx1(c))
        elif kind == 1:
            self.updateC1(c, x)
        elif kind == 2:
            self.updateC2(c, x)
        elif kind == 3:
            self.updateC3(c, x)

    def updateC1(self, c, c1):
        r = c1.findOrAdd(c)
        if r == FOUND:
            if c1.d <= 4:
                self.u = (4, Cx4(c1, c))
            else:
                self.u = (5, Cx5.fromCx1(c1, c))
        elif r == NOROOM:
            self.u = (2, Cx2(c1, c))

    def updateC2(self, c, c2):
        r = c2.findOrAdd(c)
        if r == FOUND:
            cx = Cx6()
            cx.createFrom2(c2, c)
            self.u = (6, cx)
        elif r == NOROOM:
            self.u = (3, Cx3(c2, c))

    def updateC3(self, c, c3):
        r = c3.findOrAdd(c)
        if r == FOUND:
            cx = Cx7()
            cx.createFrom3(c3, c)
            self.u = (7, cx)


# ------------------------------------------------------------------------------------------------ RangeCoder.hx
class RangeCoder:
    TOP = 0x01000000
    BOT = 0x010000

    def __init__(self):
        self.range = 0
        self.code = 0
        self.data = None
        self.pos = 0
        self.last = (0, 0, 0)       # (cumFreq, freq, total) of the last decode(): trace hook, not in the reference

    def _byte(self, i):
        v = self.data[i]
        if v is UNDEF:
            raise JsSemanticsError("range coder read past the end of the frame (code becomes NaN)")
        return v

    def DecodeBegin(self, src, pos0):
        ff = 0xFFFF
        self.range = ff * 65536
        self.range += ff
        self.data = src
        self.pos = pos0
        self.code = 0
        self.code = (self.code * 256) + self._byte(self.pos + 1)
        self.code = (self.code * 256) + self._byte(self.pos + 2)
        self.code = (self.code * 256) + self._byte(self.pos + 3)
        self.code = (self.code * 256) + self._byte(self.pos + 4)
        self.pos += 5

    def decode(self, cumFreq, freq, total_freq):
        self.last = (cumFreq, freq, total_freq)
        self.code -= cumFreq * self.range
        self.range = self.range * freq
        if self.code < 0 or self.code >= 1 << 53 or self.range >= 1 << 53:
            raise JsSemanticsError("range coder state left the exact-double domain")
        if self.range == 0:
            raise JsSemanticsError("zero range (the reference would loop for ever)")
        while self.range < RangeCoder.TOP:
            self.code = (self.code * 256) + self._byte(self.pos)
            self.pos += 1
            self.range *= 256

    def get_freq(self, total_freq):
        if total_freq == 0:
            raise JsSemanticsError("division by a zero total")
        self.range = self.range // total_freq           # Std.int(range / total_freq), exact: both < 2^53, non-negative
        if self.range == 0:
            raise JsSemanticsError("code / 0")
        return self.code // self.range

    def DecodeVal(self, cnt, maxc, step):
        totfr = cnt[maxc]
        value = self.get_freq(totfr)
        c = 0
        cumfr = 0
        cnt_c = 0
        while c < maxc:
            cnt_c = cnt[c]
            if value >= cumfr + cnt_c:
                cumfr += cnt_c
            else:
                break
            c += 1
        self.decode(cumfr, cnt_c, totfr)
        cnt[c] = cnt_c + step
        totfr += step
        if totfr > RangeCoder.BOT:
            EVENTS["rc_rescale_%d" % maxc] += 1
            totfr = 0
            for i in range(maxc):
                nc = (cnt[i] >> 1) + 1
                cnt[i] = nc
                totfr += nc
        cnt[maxc] = totfr
        return c

    def DecodeValUni(self, cnt, off, step):
        totfr = cnt[off + 16]
        value = self.get_freq(totfr)
        x = 0
        cumfr = 0
        cnt_x = 0
        while x < 16:
            cnt_x = cnt[off + x]
            if value >= cumfr + cnt_x:
                cumfr += cnt_x
            else:
                break
            x += 1
        c = x * 16
        cnt_c = 0
        while c < 256:
            cnt_c = cnt[off + c + 17]
            if value >= cumfr + cnt_c:
                cumfr += cnt_c
            else:
                break
            c += 1
        self.decode(cumfr, cnt_c, totfr)
        cnt[off + c + 17] = cnt_c + step
        cnt[off + x] = cnt_x + step
        totfr += step
        if totfr > RangeCoder.BOT:
            EVENTS["rc_rescale_uni"] += 1
            totfr = 0
            for i in range(off + 17, off + 256 + 17):
                nc = (cnt[i] >> 1) + 1
                cnt[i] = nc
                totfr += nc
            for i in range(16):
                sum_ = 0
                i16_17 = off + (i << 4) + 17
                for j in range(16):
                    sum_ += cnt[i16_17 + j]
                cnt[off + i] = sum_
        cnt[off + 16] = totfr
        return c


CXMAX = 4096
NCXMAX = 6
MSR_X = 256
MSR_Y = 256


class _LazyRows:
    """cntab = new Uint32Array(3 * 4096 * 273) (EntroCoders.hx:55): 13.4 MB of zeros in the reference; here the same
    flat index space backed by a dict of 273-word rows so that Python does not allocate 3.3 M objects.  Pure storage
    detail: reads of untouched words give 0 like a fresh typed array."""
    CNTABSZ = 273

    def __init__(self):
        self.rows = {}

    def __getitem__(self, i):
        r = self.rows.get(i // 273)
        return 0 if r is None else r[i % 273]

    def __setitem__(self, i, v):
        k = i // 273
        r = self.rows.get(k)
        if r is None:
            r = self.rows[k] = [0] * 273
        r[i % 273] = v & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ EntroCoders.hx:31-180
class EntroCoderRC:
    SC_STEP = 400
    SC_NSTEP = 400
    SC_BTSTEP = 10
    SC_BTNSTEP = 20
    SC_SXYSTEP = 100
    SC_MSTEP = 100
    SC_UNSTEP = 1000
    SC_XXSTEP = 1
    CNTABSZ = 273

    def __init__(self, trace=None):
        self.rc = RangeCoder()
        self.cntab = _LazyRows()
        self.ptypetab = [U32(7) for _ in range(NCXMAX)]
        self.ntab = [U32(257) for _ in range(NCXMAX)]
        self.xxtab = U32(257)
        self.ntab2 = U32(257)
        self.bttab = U32(6)
        self.sxytab = [U32(17) for _ in range(4)]
        self.mvtab = [U32(MSR_X * 2 + 1), U32(MSR_Y * 2 + 1)]
        self.trace = trace

    def differentConstantsFor16bbp(self):
        return True

    def preinit(self):
        # zeroes word 16 of every row: on a fresh (all-zero) table a no-op
        for k in list(self.cntab.rows):
            self.cntab.rows[k][16] = 0

    def renewI(self):
        # "fill if changed": every row whose total != 256 is reset to all-ones.  On the lazy storage: untouched rows
        # have total 0 != 256 and WOULD be filled, so materialise-on-read is replaced by: drop every row, and let
        # _row_fresh() hand out a filled row on first touch.  Same values, no 12288-row loop in Python.
        self.cntab.rows.clear()
        self.cntab_filled = True
        for ncx in range(NCXMAX):
            p = self.ntab[ncx]
            for i in range(256):
                p[i] = 1
            p[256] = 256
        for ctx in range(6):
            p = self.ptypetab[ctx]
            for i in range(6):
                p[i] = 1
            p[6] = 6
        for i in range(256):
            self.xxtab[i] = 1
            self.ntab2[i] = 1
        self.xxtab[256] = 256
        self.ntab2[256] = 256
        for i in range(5):
            self.bttab[i] = 1
        self.bttab[5] = 5
        for ctx in range(4):
            for i in range(16):
                self.sxytab[ctx][i] = 1
            self.sxytab[ctx][16] = 16
        for i in range(MSR_X * 2):
            self.mvtab[0][i] = 1
        self.mvtab[0][MSR_X * 2] = MSR_X * 2
        for i in range(MSR_Y * 2):
            self.mvtab[1][i] = 1
        self.mvtab[1][MSR_Y * 2] = MSR_Y * 2

    def _touch_row(self, cxi):
        if cxi not in self.cntab.rows:
            self.cntab.rows[cxi] = [16] * 16 + [256] + [1] * 256        # what renewI wrote (EntroCoders.hx:85-91)

    def decodeBegin(self, src, pos0):
        self.rc.DecodeBegin(src, pos0)

    def _t(self, kind, c):
        if self.trace is not None:
            cum, fr, tot = self.rc.last
            self.trace.append((kind, c, fr, cum, tot))
        return c

    def decodeClr(self, cxi):
        self._touch_row(cxi)
        return self._t(0, self.rc.DecodeValUni(self.cntab, cxi * self.CNTABSZ, self.SC_STEP))

    def decodeN(self, ptype):
        return self._t(1, self.rc.DecodeVal(self.ntab[ptype], 256, self.SC_NSTEP))

    def decodeP(self, ptype):
        return self._t(2, self.rc.DecodeVal(self.ptypetab[ptype], 6, self.SC_UNSTEP))

    def decodeX(self):
        return self._t(3, self.rc.DecodeVal(self.xxtab, 256, self.SC_XXSTEP))

    def decodeBT(self):
        return self._t(4, self.rc.DecodeVal(self.bttab, 5, self.SC_BTSTEP))

    def decodeBN(self):
        return self._t(5, self.rc.DecodeVal(self.ntab2, 256, self.SC_BTNSTEP))

    def decodeSXY(self, n):
        return self._t(6, self.rc.DecodeVal(self.sxytab[n], 16, self.SC_SXYSTEP))

    def decodeMX(self):
        return self._t(7, self.rc.DecodeVal(self.mvtab[0], MSR_X * 2, self.SC_MSTEP))

    def decodeMY(self):
        return self._t(8, self.rc.DecodeVal(self.mvtab[1], MSR_Y * 2, self.SC_MSTEP))

    def canDecodeBool(self):
        return False

    def decodeBool(self):
        return False


# ------------------------------------------------------------------------------------------------ EntroCoders.hx:182-313
class EntroCoderANS:
    def __init__(self, f0val, trace=None):
        self.myRcv = DecReceiver()
        self.cntab = [Context() for _ in range(CXMAX * 3)]
        self.ntab = [FixedSizeRansCtx(256) for _ in range(NCXMAX)]
        self.ptypetab = [FixedSizeRansCtx(6) for _ in range(6)]
        self.xxtab = FixedSizeRansCtx(256)
        self.ntab2 = FixedSizeRansCtx(256)
        self.bttab = FixedSizeRansCtx(5)
        self.sxytab = [FixedSizeRansCtx(16) for _ in range(4)]
        self.mvtab = [FixedSizeRansCtx(512) for _ in range(2)]
        Cx6.f0 = f0val
        self.rans = None
        self.nDec = 0
        self.trace = trace
        self.kinds_seen = set()

    def preinit(self):
        pass

    def differentConstantsFor16bbp(self):
        return False

    def renewI(self):
        for cx in self.cntab:
            cx.renew()
        for i in range(NCXMAX):
            self.ntab[i].renew()
        for i in range(6):
            self.ptypetab[i].renew()
        self.xxtab.renew()
        self.ntab2.renew()
        self.bttab.renew()
        for i in range(4):
            self.sxytab[i].renew()
        for i in range(2):
            self.mvtab[i].renew()

    def decodeBegin(self, src, pos0):
        oob = self.rans.oob_reads if self.rans is not None else 0
        self.rans = Rans(src, pos0)
        self.rans.oob_reads += oob
        self.nDec = 0

    def _count(self):
        self.nDec += 1
        if self.nDec == Rans.B:
            self.rans.reinit()
            self.nDec = 0

    def decodeClr(self, cxi):
        dcx = self.cntab[cxi]
        rcv = Context.rcv
        k0 = dcx.u[0]
        if dcx.decode(self.rans.decGet()):
            c = rcv.c
            self.rans.decAdvance(rcv.cumFreq, rcv.freq)
            if self.trace is not None:
                self.trace.append((0, c, rcv.freq, rcv.cumFreq, 4096))
        else:
            c = self.rans.raw()
            dcx.update(c)
            if self.trace is not None:
                self.trace.append((0, c, 0, 0, 0))
        if k0 != dcx.u[0]:
            EVENTS["kind_%d_to_%d" % (k0, dcx.u[0])] += 1
        self.kinds_seen.add((k0, dcx.u[0]))
        self._count()
        return c

    def canDecodeBool(self):
        return True

    def decodeBool(self):
        f = self.rans.decGet()
        flag = f >= Rans.PROB_SCALE >> 1
        self.rans.decAdvance((Rans.PROB_SCALE >> 1) if flag else 0, Rans.PROB_SCALE >> 1)
        if self.trace is not None:
            self.trace.append((9, 1 if flag else 0, Rans.PROB_SCALE >> 1, (Rans.PROB_SCALE >> 1) if flag else 0, 4096))
        self._count()
        return flag

    def decodeF(self, dcx, kind):
        dcx.decode(self.rans.decGet(), self.myRcv)
        self.rans.decAdvance(self.myRcv.cumFreq, self.myRcv.freq)
        if self.trace is not None:
            self.trace.append((kind, self.myRcv.c, self.myRcv.freq, self.myRcv.cumFreq, 4096))
        self._count()
        return self.myRcv.c

    def decodeN(self, ptype):
        return self.decodeF(self.ntab[ptype], 1)

    def decodeP(self, ptype):
        return self.decodeF(self.ptypetab[ptype], 2)

    def decodeX(self):
        return self.decodeF(self.xxtab, 3)

    def decodeBT(self):
        return self.decodeF(self.bttab, 4)

    def decodeBN(self):
        return self.decodeF(self.ntab2, 5)

    def decodeSXY(self, n):
        return self.decodeF(self.sxytab[n], 6)

    def decodeMX(self):
        return self.decodeF(self.mvtab[0], 7)

    def decodeMY(self):
        return self.decodeF(self.mvtab[1], 8)


ZERO_STATE, IN_PROGRESS, ERROR_OCCURED = 0, 1, 2


class FrameBuf:
    """An Int32Array of X*Y*4 ELEMENTS (Manager.hx:114-118 allocates 4x the picture), plus the Uint8Array view of its
    buffer the decoder makes (`new Uint8Array(dst.buffer)`, little endian)."""

    def __init__(self, n_pixels, fill=0):
        self.a = [fill] * (n_pixels * 4)

    def get(self, i):
        return self.a[i] if 0 <= i < len(self.a) else UNDEF

    def put(self, i, v):
        if 0 <= i < len(self.a):
            self.a[i] = 0 if v is UNDEF else _i32(v)

    def byte(self, k):
        """dstbytes[k]"""
        if k < 0 or k >= len(self.a) * 4:
            return UNDEF
        return ((self.a[k >> 2] & 0xFFFFFFFF) >> ((k & 3) * 8)) & 0xFF


def _bits(v):
    """ToInt32 of a value that may be `undefined`/NaN."""
    return 0 if v is UNDEF else v


def _pred4_channel(a, b, c):
    """(a + b - c) & 0xFF in JavaScript: any `undefined` operand makes the sum NaN, and NaN & 0xFF is 0."""
    if a is UNDEF or b is UNDEF or c is UNDEF:
        return 0
    return (a + b - c) & 0xFF


# ------------------------------------------------------------------------------------------------ ScreenPressor.hx
class ScreenPressor:
    def __init__(self, width, height, bits_per_pixel, trace=None):
        self.X = width
        self.Y = height
        self.bpp = bits_per_pixel
        self.decoder_state = ZERO_STATE
        self.SC_CXSHIFT = 0 if self.bpp == 16 else 2
        self.nbx = (self.X + 15) // 16
        self.nby = (self.Y + 15) // 16
        self.bts = [0] * (self.nbx * self.nby)
        self.decodedI = False
        self.ec = None
        self.prevFrame = None
        self.last_one_was_flat = None
        self.decodingBools = False
        self.insignificant_blocks = 0
        self.cx = 0
        self.cx1 = 0
        self.trace = trace
        self.version = 0

    def initEntro(self, version):
        if version == 2:
            self.ec = EntroCoderRC(self.trace)
        elif version == 3:
            self.ec = EntroCoderANS(64, self.trace)
            self.SC_CXSHIFT = 2
        elif version == 4:
            self.ec = EntroCoderANS(32, self.trace)
            self.SC_CXSHIFT = 2
        else:
            return False
        self.version = version
        self.decodingBools = self.ec.canDecodeBool()
        self.ec.preinit()
        return True

    def Preinit(self, insignificant_lines):
        self.insignificant_blocks = self.nbx * ((insignificant_lines + 15) // 16)

    def PreviousFrame(self):
        return self.prevFrame

    def IsKeyFrame(self, data):
        if data is None or len(data) == 0:
            return False
        b = data[0]
        return b in (0x12, 0x11, 0x22, 0x21, 0x32, 0x31)

    def RenewI(self):
        self.prevFrame = None
        if self.last_one_was_flat is not None:
            return
        self.ec.renewI()            # ec == null here is a TypeError in the reference (flat frame before any coded one)

    def DecompressI(self, srcbytes, dst):
        src = U8(list(srcbytes))
        X = self.X
        di = 0
        end = X * self.Y
        clr = 0
        lasti = di
        maskcx1, shiftcx1, shiftcx = 0xFC00, 4, 18
        ec = self.ec
        if self.decoder_state == ZERO_STATE:
            head = _num(src[0])
            version = (head >> 4) + 1
            if (head & 0xF) == 1:
                clr = 0
                if self.ec is None and self.last_one_was_flat is None:
                    raise JsSemanticsError("flat key frame before any coded key frame: ec is null (ScreenPressor.hx:114)")
                self.RenewI()
                if self.bpp == 16:
                    s1 = src[1]
                    clr16 = 0 if s1 is UNDEF else _num(src[0]) + s1 * 256      # NaN -> every `&` below gives 0
                    b = (clr16 & 0x1F) << 3
                    g = ((clr16 >> 5) & 0x1F) << 3
                    r = ((clr16 >> 10) & 0x1F) << 3
                    clr = (r << 16) + (g << 8) + b
                else:
                    b, g, r = src[1], src[2], src[3]
                    # (r << 16) + (g << 8) + b with undefined: shifts coerce to 0, a bare `+ undefined` gives NaN,
                    # and NaN stored into an Int32Array is 0
                    clr = UNDEF if b is UNDEF else (_num(r) << 16) + (_num(g) << 8) + b
                for k in range(end):
                    dst.put(k, clr)
                self.prevFrame = dst
                self.last_one_was_flat = clr
                self.decodedI = True
                return ZERO_STATE
            else:
                self.last_one_was_flat = None
            if (head & 0xF) != 2:
                return ERROR_OCCURED
            if self.ec is None:
                if not self.initEntro(version):
                    return ERROR_OCCURED
            ec = self.ec
            self.RenewI()
            ec.decodeBegin(src, 1)
            self.cx = self.cx1 = 0
            k = 0
            lasti = di
            sh = self.SC_CXSHIFT
            while k < X + 1:
                r = ec.decodeClr(self.cx + self.cx1)
                self.cx1 = (self.cx << 6) & 0xFC0
                self.cx = r >> sh
                g = ec.decodeClr(4096 + self.cx + self.cx1)
                self.cx1 = (self.cx << 6) & 0xFC0
                self.cx = g >> sh
                b = ec.decodeClr(2 * 4096 + self.cx + self.cx1)
                self.cx1 = (self.cx << 6) & 0xFC0
                self.cx = b >> sh
                n = ec.decodeN(0)
                clr = (b << 16) + (g << 8) + r
                k += n
                while n > 0:
                    n -= 1
                    dst.put(di, clr)
                    di += 1
                lasti = di - 1
        if self.bpp == 16 and ec.differentConstantsFor16bbp():
            maskcx1, shiftcx1, shiftcx = 0xFF00, 2, 16
        off = -X - 1
        ptype = 0
        sh = self.SC_CXSHIFT
        while di < end:
            ptype = ec.decodeP(ptype)
            if ptype == 0:
                r = ec.decodeClr(self.cx + self.cx1)
                self.cx1 = (self.cx << 6) & 0xFC0
                self.cx = r >> sh
                g = ec.decodeClr(4096 + self.cx + self.cx1)
                self.cx1 = (self.cx << 6) & 0xFC0
                self.cx = g >> sh
                b = ec.decodeClr(2 * 4096 + self.cx + self.cx1)
                self.cx1 = (self.cx << 6) & 0xFC0
                self.cx = b >> sh
                clr = (b << 16) + (g << 8) + r
            n = ec.decodeN(ptype)
            if ptype == 0:
                while n > 0:
                    n -= 1
                    dst.put(di, clr)
                    di += 1
                lasti = di - 1
            elif ptype == 1:
                while n > 0:
                    n -= 1
                    dst.put(di, dst.get(lasti))
                    lasti = di
                    di += 1
                clr = dst.get(lasti)
            elif ptype == 2:
                while n > 0:
                    n -= 1
                    clr = dst.get(di + off + 1)
                    dst.put(di, clr)
                    di += 1
                lasti = di - 1
            elif ptype == 4:
                while n > 0:
                    n -= 1
                    r = _pred4_channel(dst.byte(lasti * 4), dst.byte((di + off) * 4 + 4), dst.byte((di + off) * 4))
                    g = _pred4_channel(dst.byte(lasti * 4 + 1), dst.byte((di + off) * 4 + 5), dst.byte((di + off) * 4 + 1))
                    b = _pred4_channel(dst.byte(lasti * 4 + 2), dst.byte((di + off) * 4 + 6), dst.byte((di + off) * 4 + 2))
                    clr = (b << 16) + (g << 8) + r
                    dst.put(di, clr)
                    lasti = di
                    di += 1
            elif ptype == 5:
                while n > 0:
                    n -= 1
                    clr = dst.get(di + off)
                    dst.put(di, clr)
                    di += 1
                lasti = di - 1
            # ptype 3 in an I frame: no case in the switch -- nothing written, di does not move (ScreenPressor.hx:242-273)
            self.cx1 = (_bits(clr) & maskcx1) >> shiftcx1
            self.cx = _bits(clr) >> shiftcx
        self.prevFrame = dst
        self.decoder_state = ZERO_STATE
        self.decodedI = True
        return ZERO_STATE

    def DecompressP(self, srcbytes, dst):
        """-> (data_pnt, significant_changes)"""
        src = U8(list(srcbytes))
        self.last_one_was_flat = None
        if src.length == 0 or not self.decodedI:
            return self.prevFrame, False
        changes = src[0]
        if changes == 0:
            return self.prevFrame, False
        ec = self.ec
        maskcx1, shiftcx1, shiftcx = 0xFC00, 4, 18
        if ec.differentConstantsFor16bbp() and self.bpp == 16:
            maskcx1, shiftcx1, shiftcx = 0xFF00, 2, 16
        ec.decodeBegin(src, 1)
        t = ec.decodeX()
        xx1 = ec.decodeX()
        xx1 = (xx1 << 8) + t
        t = ec.decodeX()
        xx2 = ec.decodeX()
        xx2 = (xx2 << 8) + t
        bts = self.bts
        nbts = len(bts)
        for i in range(nbts):
            bts[i] = 0
        x = xx1
        while x <= xx2:
            block_type = ec.decodeBT()
            n = ec.decodeBN()
            for _ in range(n):
                if 0 <= x < nbts:           # Int32Array: out-of-range stores are dropped
                    bts[x] = block_type
                x += 1
        signif = False
        for i in range(self.insignificant_blocks, nbts):
            if bts[i] > 0:
                signif = True
                break
        X, Y = self.X, self.Y
        stride = X
        clr = 0
        off = -X - 1
        self.cx = self.cx1 = 0
        lastmx = lastmy = 0
        prevFrame = self.prevFrame
        sh = self.SC_CXSHIFT
        for by in range(self.nby):
            for bx in range(self.nbx):
                y16 = by * 16
                x16 = bx * 16
                x1 = x16
                x2 = x16 + 16
                y1 = y16
                y2 = y16 + 16
                if x2 > X:
                    x2 = X
                if y2 > Y:
                    y2 = Y
                bi = by * self.nbx + bx
                if bts[bi] > 0:
                    if ((bts[bi] - 1) & 1) > 0:
                        for y in range(y1, y2):
                            i = y * stride + x1
                            for xo in range(x2 - x1):
                                dst.put(i + xo, prevFrame.get(i + xo))
                        x1 = ec.decodeSXY(0) + x16
                        y1 = ec.decodeSXY(1) + y16
                        x2 = ec.decodeSXY(2) + x16 + 1
                        y2 = ec.decodeSXY(3) + y16 + 1
                    if ((bts[bi] - 1) & 2) > 0:
                        if self.decodingBools and ec.decodeBool():
                            mx = lastmx
                            my = lastmy
                        else:
                            mx = ec.decodeMX() - MSR_X
                            my = ec.decodeMY() - MSR_Y
                        lastmx = mx
                        lastmy = my
                        for y in range(y1, y2):
                            i = y * stride + x1
                            j = (y + my) * stride + (x1 + mx)
                            for xo in range(x2 - x1):
                                dst.put(i + xo, prevFrame.get(j + xo))
                    else:
                        x = x1
                        y = y1
                        ptype = 0
                        while y < y2:
                            i = y * stride + x
                            di = i
                            lastptype = ptype
                            ptype = ec.decodeP(lastptype)
                            if ptype == 0:
                                r = ec.decodeClr(self.cx + self.cx1)
                                self.cx1 = (self.cx << 6) & 0xFC0
                                self.cx = r >> sh
                                g = ec.decodeClr(4096 + self.cx + self.cx1)
                                self.cx1 = (self.cx << 6) & 0xFC0
                                self.cx = g >> sh
                                b = ec.decodeClr(2 * 4096 + self.cx + self.cx1)
                                self.cx1 = (self.cx << 6) & 0xFC0
                                self.cx = b >> sh
                                clr = (b << 16) + (g << 8) + r
                            n = ec.decodeN(ptype)
                            for _ in range(n):
                                if ptype == 1:
                                    clr = dst.get(di - 1)
                                elif ptype == 2:
                                    clr = dst.get(di + off + 1)
                                elif ptype == 3:
                                    clr = prevFrame.get(i)
                                elif ptype == 4:
                                    r = _pred4_channel(dst.byte((di - 1) * 4), dst.byte((di + off) * 4 + 4), dst.byte((di + off) * 4))
                                    g = _pred4_channel(dst.byte((di - 1) * 4 + 1), dst.byte((di + off) * 4 + 5), dst.byte((di + off) * 4 + 1))
                                    b = _pred4_channel(dst.byte((di - 1) * 4 + 2), dst.byte((di + off) * 4 + 6), dst.byte((di + off) * 4 + 2))
                                    clr = (b << 16) + (g << 8) + r
                                elif ptype == 5:
                                    clr = dst.get(di + off)
                                dst.put(di, clr)
                                x += 1
                                if x >= x2:
                                    x = x1
                                    y += 1
                                    i = y * stride + x
                                    di = i
                                else:
                                    i += 1
                                    di += 1
                            self.cx1 = (_bits(clr) & maskcx1) >> shiftcx1
                            self.cx = _bits(clr) >> shiftcx
                else:
                    for y in range(y1, y2):
                        i = y * stride + x1
                        for xo in range(x2 - x1):
                            dst.put(i + xo, prevFrame.get(i + xo))
        self.prevFrame = dst
        return self.prevFrame, signif


def decode_stream(width, height, bpp, frames, insignificant_lines=0, trace=False, stale=None):
    """Drive one ScreenPressor instance over `frames` (list of bytes) the way Manager.worker does
    (Manager.hx:458-525): IsKeyFrame -> DecompressI, else DecompressP; every call gets a `dst` that is neither the
    codec's prevFrame nor (here) ever reused.  `stale`: what a fresh dst holds before a P frame is decoded into it --
    None = a copy of the previous picture (the defined behaviour of this repository), an int = that
    value in every element (to expose what depends on stale ring-buffer contents).
    Returns (pictures [n][X*Y] lists, changed [n], significant [n], traces [n] (lists of tuples) or None, info)."""
    tr = [] if trace else None
    sp = ScreenPressor(width, height, bpp, trace=tr)
    sp.Preinit(insignificant_lines)
    n_px = width * height
    pics, changed, signif, traces = [], [], [], []
    for fb in frames:
        if tr is not None:
            del tr[:]
        if sp.IsKeyFrame(fb):
            dst = FrameBuf(n_px)
            st = sp.DecompressI(fb, dst)
            if st != ZERO_STATE:
                pics.append(None)
                changed.append(False)
                signif.append(False)
                traces.append(list(tr) if tr is not None else None)
                continue
            pics.append(dst.a[:n_px])
            changed.append(True)
            signif.append(False)        # DecompressI has no significant_changes; the batch convention reports 0
        else:
            prev = sp.PreviousFrame()
            dst = FrameBuf(n_px)
            if prev is not None:
                if stale is None:
                    dst.a[:] = prev.a
                else:
                    dst.a[:] = [stale] * len(dst.a)
            res, sg = sp.DecompressP(fb, dst)
            pics.append(None if res is None else res.a[:n_px])
            changed.append(res is dst)
            signif.append(bool(sg))
        traces.append(list(tr) if tr is not None else None)
    info = {"version": sp.version}
    if isinstance(sp.ec, EntroCoderANS):
        info["kind_transitions"] = sorted(sp.ec.kinds_seen)
        info["oob_reads"] = sp.ec.rans.oob_reads if sp.ec.rans is not None else 0
    return pics, changed, signif, traces, info
