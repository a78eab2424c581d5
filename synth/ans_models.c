/*
 * ans_models.c -- CPU restatement of the adaptive colour-context models of reference src/ANS.hx (SymbList/Cx1-3
 * :155-208, SmallContext/Cx4/Cx5 :210-392, Cx6 :394-704, Cx7 :706-772, Context :785-860, Sorter :862-872); the
 * fixed-size table (FixedSizeRansCtx :54-145) is inline in ans_models.h.
 *
 * Lives with the synthetic encoder because a rANS ENCODER must run the decoder's models symbol by symbol to know
 * every interval.  The repo's CPU checker compiles this same file for its decoder (the
 * dependency points from the checker to this file, never the other way); the CUDA decoder (csrc/sp_ans.cu) is an independent implementation and is checked against it.
 * Not part of libjsplayer_cuda.
 */
#include "ans_models.h"
#include <stdlib.h>

/* test hook: how often each context transition ran (coverage of the synthetic streams; racy under threads, only
 * read by single-threaded tests) */
static long g_trans[8];
enum { TR_4FROM1, TR_5FROM1, TR_5FROM4, TR_6FROM5, TR_6FROM2, TR_7FROM3, TR_7FROM6, TR_6GROW };
void jsp_ans_transitions(long out[8], int reset) { for (int i = 0; i < 8; i++) { out[i] = g_trans[i]; if (reset) g_trans[i] = 0; } }

/* ---------------------------------------------------------------- colour contexts ---- */
static void insort(uint8_t *a, int n)                     /* Sorter.insort, ANS.hx:862-872 */
{
    for (int i = 1; i < n; i++) {
        int j = i;
        while (j > 0 && a[j - 1] > a[j]) { uint8_t t = a[j]; a[j] = a[j - 1]; a[j - 1] = t; j--; }
    }
}

enum { FOUND, ADDED, NOROOM };
static int find_or_add(color_ctx *x, int c, int cap)      /* SymbList.findOrAdd, :163-171 */
{
    for (int i = 0; i < x->d; i++) if (x->symb[i] == (uint8_t)c && c >= 0 && c < 256) return FOUND;
    if (x->d < cap) { x->symb[x->d] = (uint8_t)c; x->d++; return ADDED; }
    return NOROOM;
}

void cctx_renew(color_ctx *x) { x->kind = CXK_NONE; }
void cctx_free(color_ctx *x) { free(x->c7); x->c7 = NULL; }

/* SmallContext.create, :226-238 (from the Cx1 list held in x->symb) */
static void sc_create(color_ctx *x, int S, int c)
{
    x->S = S; x->maxpos = 0;
    memset(x->sc_symbols, 0, sizeof x->sc_symbols); memset(x->sc_freqs, 0, sizeof x->sc_freqs);
    insort(x->symb, x->d);
    for (int i = 0; i < x->d; i++) {
        x->sc_symbols[i] = x->symb[i];
        if (x->sc_symbols[i] == c) { x->sc_freqs[i] = 100; x->maxpos = i; } else x->sc_freqs[i] = 50;
    }
}
static void sc_rescale(color_ctx *x, int *totFr)          /* :254-261 */
{
    int s = 256 - x->d;
    for (int i = 0; i < x->d; i++) { x->sc_freqs[i] = (uint16_t)(x->sc_freqs[i] - (x->sc_freqs[i] >> 1)); s += x->sc_freqs[i]; }
    *totFr = s;
}
static int sc_add(color_ctx *x, int pos, int c, int *totFr)   /* addSymb, :240-252 */
{
    if (x->d == x->S) return 0;
    for (int i = x->d - 1; i >= pos; i--) { x->sc_symbols[i + 1] = x->sc_symbols[i]; x->sc_freqs[i + 1] = x->sc_freqs[i]; }
    x->sc_symbols[pos] = (uint8_t)c; x->sc_freqs[pos] = 50; x->d++;
    if (x->maxpos >= pos) x->maxpos++;
    *totFr += 50;
    if (*totFr + 50 > ANS_PROB_SCALE) sc_rescale(x, totFr);
    return 1;
}
/* SmallContext.decodeSC, :263-309 */
static int sc_decode(color_ctx *x, int someFreq, dec_receiver *r, int totFr0, int *totFr_out)
{
    int totFr = totFr0, shift = 0, tot = totFr0;
    while (tot <= ANS_PROB_SCALE / 2 && tot > 0) { tot <<= 1; shift++; }
    someFreq >>= shift;
    const int bonus = (ANS_PROB_SCALE - tot) >> shift;
    const uint16_t maxFreq = x->sc_freqs[x->maxpos];
    x->sc_freqs[x->maxpos] = (uint16_t)(maxFreq + bonus);
    int cumFr = 0, lastSymb = 0, pos = 0, res;
    while (pos < x->d) {
        const int s = x->sc_symbols[pos];
        const int startFr = cumFr + s - lastSymb;
        if (someFreq < startFr) {
            r->c = someFreq - cumFr + lastSymb;
            cumFr = someFreq;
            r->cumFreq = cumFr << shift; r->freq = 1 << shift;
            x->sc_freqs[x->maxpos] = maxFreq;
            res = sc_add(x, pos, r->c, &totFr);
            *totFr_out = totFr;
            return res;
        }
        const int fr = x->sc_freqs[pos];
        if (startFr + fr > someFreq) {
            r->c = s;
            cumFr += r->c - lastSymb;
            r->cumFreq = cumFr << shift; r->freq = fr << shift;
            x->sc_freqs[x->maxpos] = maxFreq;
            x->sc_freqs[pos] = (uint16_t)(x->sc_freqs[pos] + 50); totFr += 50;
            if (pos != x->maxpos && x->sc_freqs[pos] > x->sc_freqs[x->maxpos]) x->maxpos = pos;
            if (totFr + 50 > ANS_PROB_SCALE) sc_rescale(x, &totFr);
            *totFr_out = totFr;
            return 1;
        }
        cumFr += s - lastSymb + fr;
        lastSymb = s + 1;
        pos++;
    }
    x->sc_freqs[x->maxpos] = maxFreq;
    r->c = lastSymb + someFreq - cumFr;
    r->cumFreq = someFreq << shift; r->freq = 1 << shift;
    res = sc_add(x, pos, r->c, &totFr);
    *totFr_out = totFr;
    return res;
}
static int c4_total(const color_ctx *x) { return x->sc_freqs[0] + x->sc_freqs[1] + x->sc_freqs[2] + x->sc_freqs[3] + 256 - x->d; }   /* :320 */
static void c5_calcsum(color_ctx *x) { int t = 256 - x->d; for (int i = 0; i < x->d; i++) t += x->sc_freqs[i]; x->cntsum5 = t; }       /* :374-378 */

/* Cx5.createFrom4, :350-372 (in place: x is the Cx4) */
static void c5_from4(color_ctx *x, int c)
{
    uint8_t os[4]; uint16_t of[4]; const int dd = x->d;
    memcpy(os, x->sc_symbols, 4); memcpy(of, x->sc_freqs, 8);
    memset(x->sc_symbols, 0, sizeof x->sc_symbols); memset(x->sc_freqs, 0, sizeof x->sc_freqs);
    x->S = 16; x->maxpos = 0;                              /* a new Cx5: maxpos starts at 0 (:223) */
    int i = 0, totFr = 0;
    while (i < dd && os[i] < c) { x->sc_symbols[i] = os[i]; totFr += (x->sc_freqs[i] = of[i]); i++; }
    int j = i;
    x->sc_symbols[j] = (uint8_t)c; totFr += (x->sc_freqs[j] = 50); j++;
    while (i < dd) { x->sc_symbols[j] = os[i]; totFr += (x->sc_freqs[j] = of[i]); i++; j++; }
    x->d = dd + 1;
    if (totFr > ANS_PROB_SCALE) { int t; sc_rescale(x, &t); }
    c5_calcsum(x);
    x->kind = CXK_5;
    g_trans[TR_5FROM4]++;
}

/* ---- Cx6 ---- */
static inline void c6_set(color_ctx *x, int i, int fr, int cf) { x->c6_freqs[i * 2] = (uint16_t)fr; x->c6_freqs[i * 2 + 1] = (uint16_t)cf; }
static void c6_init(color_ctx *x, int S)
{
    x->S6 = S;
    memset(x->c6_symbols, 0, sizeof x->c6_symbols); memset(x->c6_freqs, 0, sizeof x->c6_freqs); memset(x->c6_cnts, 0, sizeof x->c6_cnts);
}
static void c6_calcsum(color_ctx *x)                       /* :571-578 */
{
    const int shft = x->fshift > 0 ? x->fshift - 1 : 0;
    int sum = (256 - x->d) << shft;
    for (int i = 0; i < x->S6; i++) sum += x->c6_cnts[i];
    x->c6_cnts[x->S6] = (uint16_t)sum;
}
static void c6_rescale(color_ctx *x)                       /* rescaleDec, :580-604 */
{
    uint16_t _cnts[256], _freqs[512];
    const int sh = x->fshift > 0 ? x->fshift - 1 : 0;
    const int c0 = 1 << sh;
    for (int i = 0; i < 256; i++) _cnts[i] = (uint16_t)c0;
    for (int i = 0; i < x->d; i++) _cnts[x->c6_symbols[i]] = x->c6_cnts[i];
    int cumFr = 0;
    for (int i = 0; i < 256; i++) { _freqs[i * 2] = _cnts[i]; _freqs[i * 2 + 1] = (uint16_t)cumFr; cumFr += _cnts[i]; }
    if (x->fshift > 0) x->fshift--;
    const int shft = x->fshift > 0 ? x->fshift - 1 : 0;
    int cntsum = (256 - x->d) << shft;
    for (int i = 0; i < x->d; i++) {
        x->c6_cnts[i] = (uint16_t)(x->c6_cnts[i] - (x->c6_cnts[i] >> 1));
        cntsum += x->c6_cnts[i];
        const int idx = x->c6_symbols[i];
        c6_set(x, i, _freqs[idx * 2], _freqs[idx * 2 + 1]);
    }
    x->c6_cnts[x->S6] = (uint16_t)cntsum;
}
static void c6_incr(color_ctx *x, int pos)                 /* incrCntDec, :680-696 */
{
    const int step = 25 << x->fshift, S = x->S6;
    x->c6_cnts[pos] = (uint16_t)(x->c6_cnts[pos] + step);
    x->c6_cnts[S] = (uint16_t)(x->c6_cnts[S] + step);
    if (pos > 0 && x->c6_cnts[pos] > x->c6_cnts[pos - 1]) {
        uint16_t tc = x->c6_cnts[pos]; x->c6_cnts[pos] = x->c6_cnts[pos - 1]; x->c6_cnts[pos - 1] = tc;
        uint16_t fp = x->c6_freqs[pos * 2], cfp = x->c6_freqs[pos * 2 + 1];
        c6_set(x, pos, x->c6_freqs[(pos - 1) * 2], x->c6_freqs[(pos - 1) * 2 + 1]);
        c6_set(x, pos - 1, fp, cfp);
        uint8_t ts = x->c6_symbols[pos]; x->c6_symbols[pos] = x->c6_symbols[pos - 1]; x->c6_symbols[pos - 1] = ts;
    }
    if (x->c6_cnts[S] + step > ANS_PROB_SCALE) c6_rescale(x);
}
/* Cx6.createFrom5, :431-505 (in place: x is the Cx5; c did not fit) */
static void c6_from5(color_ctx *x, int c)
{
    uint8_t os[16]; uint16_t of[16]; const int oldd = x->d;
    memcpy(os, x->sc_symbols, 16); memcpy(of, x->sc_freqs, 32);
    c6_init(x, 32);
    int totFr = 256 - oldd;
    for (int i = 0; i < oldd; i++) totFr += of[i];
    int shift = 0, tot = totFr;
    while (tot <= ANS_PROB_SCALE / 2 && tot > 0) { tot <<= 1; shift++; }
    int cumFr = 0, lastSymb = 0;
    for (int pos = 0; pos < oldd; pos++) {
        const int s = os[pos];
        cumFr += s - lastSymb;
        const int cfr = of[pos], fr = cfr << shift;
        c6_set(x, pos, fr, cumFr << shift);
        x->c6_cnts[pos] = (uint16_t)(fr - (fr >> 1));
        x->c6_symbols[pos] = (uint8_t)s;
        cumFr += cfr;
        lastSymb = s + 1;
    }
    x->fshift = shift;
    const int fr_freq = 1 << shift; int fr_cum = 0;
    if (c > 0) {
        int lowerSym = -1, lfreq = 0, lcum = 0;
        for (int i = 0; i < oldd; i++) {
            const int s = x->c6_symbols[i];
            if (s > lowerSym && s < c) { lowerSym = s; lfreq = x->c6_freqs[i * 2]; lcum = x->c6_freqs[i * 2 + 1]; }
        }
        fr_cum = lfreq > 0 ? lcum + lfreq + ((c - lowerSym - 1) << shift) : (c << shift);
    }
    c6_set(x, oldd, fr_freq, fr_cum);
    x->c6_cnts[oldd] = (uint16_t)(fr_freq - (fr_freq >> 1));
    x->c6_symbols[oldd] = (uint8_t)c;
    x->d = oldd + 1;
    const int step = 25 << shift, S = 32;
    x->c6_cnts[oldd] = (uint16_t)(x->c6_cnts[oldd] + step);
    x->c6_cnts[S] = (uint16_t)(x->c6_cnts[S] + step);
    if (x->c6_cnts[S] + step > ANS_PROB_SCALE) c6_rescale(x);
    c6_calcsum(x);
    for (int i = 0; i < x->d - 1; i++)                     /* sort by freqs, descending */
        for (int j = i + 1; j < x->d; j++) {
            const uint16_t fj = x->c6_freqs[j * 2], fi = x->c6_freqs[i * 2];
            if (fj > fi) {
                const uint16_t cfi = x->c6_freqs[i * 2 + 1], cfj = x->c6_freqs[j * 2 + 1];
                c6_set(x, i, fj, cfj); c6_set(x, j, fi, cfi);
                uint16_t tc = x->c6_cnts[i]; x->c6_cnts[i] = x->c6_cnts[j]; x->c6_cnts[j] = tc;
                uint8_t ts = x->c6_symbols[i]; x->c6_symbols[i] = x->c6_symbols[j]; x->c6_symbols[j] = ts;
            }
        }
    x->kind = CXK_6;
    g_trans[TR_6FROM5]++;
}
/* Cx6.createFrom2, :507-555 (x holds the Cx2 list; c was met the second time) */
static void c6_from2(color_ctx *x, int c, int f0)
{
    const int oldd = x->d;
    c6_init(x, oldd <= 32 ? 32 : 64);
    int totFr = 256 - oldd + oldd * f0 + f0;
    int shift = 0, tot = totFr;
    while (tot <= ANS_PROB_SCALE / 2 && tot > 0) { tot <<= 1; shift++; }
    int cumFr = 0, lastSymb = 0, newSymbPos = 0;
    insort(x->symb, oldd);
    for (int pos = 0; pos < oldd; pos++) {
        const int s = x->symb[pos];
        cumFr += s - lastSymb;
        int cfr;
        if (s == c) { newSymbPos = pos; cfr = f0 * 2; } else cfr = f0;
        const int fr = cfr << shift;
        c6_set(x, pos, fr, cumFr << shift);
        x->c6_symbols[pos] = (uint8_t)s;
        x->c6_cnts[pos] = (uint16_t)(fr - (fr >> 1));
        cumFr += cfr;
        lastSymb = s + 1;
    }
    x->d = oldd; x->fshift = shift;
    c6_calcsum(x);
    if (newSymbPos > 0) {
        const uint16_t fr0 = x->c6_freqs[0], cf0 = x->c6_freqs[1], frc = x->c6_freqs[newSymbPos * 2], cfc = x->c6_freqs[newSymbPos * 2 + 1];
        c6_set(x, 0, frc, cfc); c6_set(x, newSymbPos, fr0, cf0);
        const uint8_t sym0 = x->c6_symbols[0]; const uint16_t cnt0 = x->c6_cnts[0], cntc = x->c6_cnts[newSymbPos];
        x->c6_cnts[0] = cntc; x->c6_cnts[newSymbPos] = cnt0;
        x->c6_symbols[0] = (uint8_t)c; x->c6_symbols[newSymbPos] = sym0;
    }
    x->kind = CXK_6;
    g_trans[TR_6FROM2]++;
}
static int c6_add(color_ctx *x, int c, int freq, int cum)  /* addDec, :652-661 */
{
    if (x->d >= 40 || x->d >= x->S6) return -1;
    const int pos = x->d;
    x->c6_symbols[pos] = (uint8_t)c; c6_set(x, pos, freq, cum);
    x->c6_cnts[pos] = (uint16_t)(freq - (freq >> 1));
    x->d++;
    return pos;
}
static void c6_grow(color_ctx *x)                          /* growDec, :663-678: 32 -> 64 slots */
{
    const uint16_t sum = x->c6_cnts[x->S6];
    x->c6_cnts[x->S6] = 0;
    for (int i = x->d; i < 64; i++) { x->c6_symbols[i] = 0; x->c6_cnts[i] = 0; x->c6_freqs[i * 2] = 0; x->c6_freqs[i * 2 + 1] = 0; }
    x->S6 *= 2;
    x->c6_cnts[x->S6] = sum;
    g_trans[TR_6GROW]++;
}
/* Cx6.decode, :606-650; 0 = the context must be upgraded to Cx7 (r->c is the symbol) */
static int c6_decode(color_ctx *x, int someFreq, dec_receiver *r)
{
    int lfreq = 0, lcum = 0, lowerSym = 0;
    for (int i = 0; i < x->d; i++) {
        const int cf = x->c6_freqs[i * 2 + 1];
        if (cf <= someFreq) {
            const int fr = x->c6_freqs[i * 2];
            if (cf + fr > someFreq) { r->c = x->c6_symbols[i]; r->freq = fr; r->cumFreq = cf; c6_incr(x, i); return 1; }
            if (cf >= lcum) { lfreq = fr; lcum = cf; lowerSym = x->c6_symbols[i]; }
        }
    }
    const int fr_freq = 1 << x->fshift; int fr_cum, c;
    if (lfreq > 0) {
        const int cumFr = lcum + lfreq;
        const int xx = (someFreq - cumFr) >> x->fshift;
        c = xx + lowerSym + 1;
        fr_cum = lcum + lfreq + (xx << x->fshift);
    } else { c = someFreq >> x->fshift; fr_cum = c << x->fshift; }
    r->freq = fr_freq; r->cumFreq = fr_cum; r->c = c;
    int p = c6_add(x, c, fr_freq, fr_cum);
    if (p < 0) {
        if (x->S6 == 64) return 0;
        c6_grow(x);
        p = c6_add(x, c, fr_freq, fr_cum);
        if (p < 0) return 0;
    }
    c6_incr(x, p);
    return 1;
}

/* ---- Cx7 ---- */
static fixed_ctx *c7_alloc(color_ctx *x) { if (!x->c7) x->c7 = (fixed_ctx *)malloc(sizeof(fixed_ctx)); fx_init(x->c7, 256); return x->c7; }
static void c7_from3(color_ctx *x, int c)                  /* :711-739 */
{
    fixed_ctx *t = c7_alloc(x);
    for (int i = 0; i < 256; i++) { t->freqs[i * 2] = 1; t->cnts[i] = 1; }
    const int d = x->d;
    const int f0 = (ANS_PROB_SCALE - (256 - d)) / (d + 1);
    const int c0 = f0 - (f0 >> 1);
    for (int i = 0; i < d; i++) { const int s = x->symb[i]; t->freqs[s * 2] = (uint16_t)f0; t->cnts[s] = (uint16_t)c0; }
    t->freqs[c * 2] = (uint16_t)(t->freqs[c * 2] + f0);
    t->cnts[c] = (uint16_t)(t->cnts[c] + FX_STEP);
    t->cntsum = 0; int cf = 0;
    for (int i = 0; i < 256; i++) {
        t->cntsum += t->cnts[i];
        t->freqs[i * 2 + 1] = (uint16_t)cf;
        const int fr = t->freqs[i * 2];
        fx_fill_dec(t, i, cf, fr);
        cf += fr;
    }
    x->kind = CXK_7;
    g_trans[TR_7FROM3]++;
}
static void c7_from6(color_ctx *x)                         /* :741-771 */
{
    fixed_ctx *t = c7_alloc(x);
    const int S = x->S6;
    t->cntsum = x->c6_cnts[S];
    for (int i = 0; i < S; i++) if (x->c6_cnts[i] > 0) {
        const int s = x->c6_symbols[i];
        fx_set(t, s, x->c6_freqs[i * 2], x->c6_freqs[i * 2 + 1]);
        t->cnts[s] = x->c6_cnts[i];
    }
    const int funmet = 1 << x->fshift, cntUnmet = funmet - (funmet >> 1);
    int cumFr = 0;
    for (int i = 0; i < 256; i++) {
        int fr;
        if (t->freqs[i * 2] > 0) fr = t->freqs[i * 2];
        else { fx_set(t, i, funmet, cumFr); t->cnts[i] = (uint16_t)cntUnmet; fr = funmet; }
        fx_fill_dec(t, i, cumFr, fr);
        cumFr += fr;
    }
    x->kind = CXK_7;
    g_trans[TR_7FROM6]++;
}

int cctx_decode(color_ctx *x, int someFreq, dec_receiver *r, int f0)   /* Context.decode, :795-810 */
{
    (void)f0;
    int tf;
    switch (x->kind) {
    case CXK_6: if (!c6_decode(x, someFreq, r)) c7_from6(x); return 1;
    case CXK_7: fx_decode(x->c7, someFreq, r); return 1;
    case CXK_4: if (!sc_decode(x, someFreq, r, c4_total(x), &tf)) c5_from4(x, r->c); return 1;
    case CXK_5: { const int ok = sc_decode(x, someFreq, r, x->cntsum5, &tf); x->cntsum5 = tf; if (!ok) c6_from5(x, r->c); return 1; }
    default: return 0;
    }
}

void cctx_update(color_ctx *x, int c, int f0)              /* Context.update, :812-859 */
{
    switch (x->kind) {
    case CXK_NONE: x->kind = CXK_1; x->d = 1; x->symb[0] = (uint8_t)c; break;            /* new Cx1(c), :179-186 */
    case CXK_1:
        switch (find_or_add(x, c, 14)) {
        case FOUND:
            if (x->d <= 4) { sc_create(x, 4, c); x->kind = CXK_4; g_trans[TR_4FROM1]++; }
            else { sc_create(x, 16, c); c5_calcsum(x); x->kind = CXK_5; g_trans[TR_5FROM1]++; }                 /* Cx5.fromCx1, :337-342 */
            break;
        case NOROOM: x->symb[x->d] = (uint8_t)c; x->d++; x->kind = CXK_2; break;          /* new Cx2(c1,c), :188-197 */
        default: break;
        }
        break;
    case CXK_2:
        switch (find_or_add(x, c, 64)) {
        case FOUND: c6_from2(x, c, f0); break;
        case NOROOM: x->symb[x->d] = (uint8_t)c; x->d++; x->kind = CXK_3; break;          /* new Cx3(c2,c), :199-208 */
        default: break;
        }
        break;
    case CXK_3:
        if (find_or_add(x, c, 256) == FOUND) c7_from3(x, c);
        break;
    default: break;                                        /* "unexpected kind in Context.update" */
    }
}

/* encoder side: the interval the decoder computes for symbol c in the current state (read-only) */
int cctx_interval(const color_ctx *x, int c, int *freq, int *cum)
{
    switch (x->kind) {
    case CXK_7: *freq = x->c7->freqs[c * 2]; *cum = x->c7->freqs[c * 2 + 1]; break;
    case CXK_4: case CXK_5: {
        const int totFr0 = x->kind == CXK_4 ? c4_total(x) : x->cntsum5;
        int shift = 0, tot = totFr0;
        while (tot <= ANS_PROB_SCALE / 2 && tot > 0) { tot <<= 1; shift++; }
        const int bonus = (ANS_PROB_SCALE - tot) >> shift;
        int cumFr = 0, lastSymb = 0, pos = 0, done = 0;
        while (pos < x->d) {
            const int s = x->sc_symbols[pos];
            if (c < s) { *cum = (cumFr + c - lastSymb) << shift; *freq = 1 << shift; done = 1; break; }
            const int fr = x->sc_freqs[pos] + (pos == x->maxpos ? bonus : 0);
            if (c == s) { *cum = (cumFr + s - lastSymb) << shift; *freq = fr << shift; done = 1; break; }
            cumFr += s - lastSymb + fr; lastSymb = s + 1; pos++;
        }
        if (!done) { *cum = (cumFr + c - lastSymb) << shift; *freq = 1 << shift; }
        break;
    }
    case CXK_6: {
        int lowerSym = -1, lfreq = 0, lcum = 0, found = 0;
        for (int i = 0; i < x->d; i++) {
            const int s = x->c6_symbols[i];
            if (s == c) { *freq = x->c6_freqs[i * 2]; *cum = x->c6_freqs[i * 2 + 1]; found = 1; break; }
            if (s < c && s > lowerSym) { lowerSym = s; lfreq = x->c6_freqs[i * 2]; lcum = x->c6_freqs[i * 2 + 1]; }
        }
        if (!found) {
            *freq = 1 << x->fshift;
            *cum = lowerSym >= 0 ? lcum + lfreq + ((c - lowerSym - 1) << x->fshift) : (c << x->fshift);
        }
        break;
    }
    default: return 0;
    }
    if (*freq <= 0 || *cum < 0 || *cum + *freq > ANS_PROB_SCALE) return -1;
    return 1;
}
