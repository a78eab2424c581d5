/*
 * rangecoder_oracle.c -- CPU restatement of reference src/RangeCoder.hx (whole file) and of
 * EntroCoderRC (src/EntroCoders.hx:31-180).  TEST INFRASTRUCTURE ONLY.
 *
 * JavaScript numbers made explicit: `range` and `code` are doubles holding non-negative integers.  On a
 * valid stream code < range <= 2^32-1 always, so uint32/uint64 integer arithmetic is exact
 * (RangeCoder.hx:36-49; `Std.int(a/b)` = floor for these operands).  A read past the end of the data
 * yields `undefined`, which turns `code` into NaN for good; `Std.int(NaN / range)` is 0 from then on
 * (RangeCoder.hx:41,48).  That state is kept as `poisoned`; asking for another symbol in it is reported
 * through failed() (defined behaviour; the reference would go on decoding symbol 0 for ever).
 *
 * After the first failed symbol of a frame no further symbol is decoded (every later call returns 0 and touches
 * nothing): a failed symbol leaves `range` un-normalised, and what JavaScript's doubles would compute from there
 * (code >= 2^32, division by a zero range) is not something a 32-bit decoder should have to reproduce.
 */
#include "sp_entro.h"
#include <stdlib.h>
#include <string.h>

#define TOP 0x01000000u   /* RangeCoder.hx:12 */
#define BOT 0x010000u     /* RangeCoder.hx:13 */

typedef struct {
    uint32_t range;
    uint64_t code;
    const uint8_t *data;
    int len, pos;
    int poisoned, failed;
    uint32_t last_cum, last_freq, last_tot;               /* the interval of the last decoded symbol (trace hook) */
} rangecoder;

static inline void rc_byte(rangecoder *rc)
{
    if (rc->pos >= 0 && rc->pos < rc->len) rc->code = (rc->code << 8) + rc->data[rc->pos];
    else rc->poisoned = 1;
    rc->pos++;
}

/* RangeCoder.hx:19-34 */
static void rc_begin(rangecoder *rc, const uint8_t *src, int len, int pos0)
{
    rc->code = 0; rc->range = 0xFFFFFFFFu; rc->data = src; rc->len = len; rc->poisoned = 0;
    rc->failed = 0;                                       /* the failure report is per frame */
    rc->pos = pos0 + 1;
    rc_byte(rc); rc_byte(rc); rc_byte(rc); rc_byte(rc);
    /* pos is now pos0 + 5 */
}

/* RangeCoder.hx:45-49 */
/* test hook: entropy-coded symbols decoded so far by the calling thread (both coders).  Thread-local: a shared
 * counter would make the multi-threaded CPU baseline fight over one cache line. */
__thread unsigned long long g_ora_symbols = 0;
/* test hook (single-threaded callers only): when set, every entropy-coded symbol both coders hand to the frame loop is
 * appended as 5 ints: call kind (0 clr, 1 N, 2 P, 3 X, 4 BT, 5 BN, 6 SXY, 7 MX, 8 MY, 9 bool), symbol, freq, cumFreq,
 * total (range coder: the table total; rANS: 4096, or freq = cumFreq = total = 0 for a raw byte).  Compared against the
 * independent Python reading (oracle/sp_naive.py) by tests/test_sp_second_reading.py. */
int32_t *g_ora_sym_trace = 0; long g_ora_sym_trace_cap = 0, g_ora_sym_trace_n = 0;
void ora_sym_trace(int32_t *buf, long cap_symbols) { g_ora_sym_trace = buf; g_ora_sym_trace_cap = cap_symbols; g_ora_sym_trace_n = 0; }
long ora_sym_trace_count(void) { return g_ora_sym_trace_n; }
void ora_sym_trace_put(int kind, int c, int freq, int cum, int tot)
{
    if (!g_ora_sym_trace) return;
    if (g_ora_sym_trace_n < g_ora_sym_trace_cap) {
        int32_t *p = g_ora_sym_trace + 5 * g_ora_sym_trace_n;
        p[0] = kind; p[1] = c; p[2] = freq; p[3] = cum; p[4] = tot;
    }
    g_ora_sym_trace_n++;
}
unsigned long long ora_symbol_count(int reset) { unsigned long long v = g_ora_symbols; if (reset) g_ora_symbols = 0; return v; }

/* Defined behaviour: a failed frame decodes nothing more -- every later call returns 0 and leaves the models alone.
 * Asking for a symbol after the data has run out (poisoned) is such a failure.  Calls are still counted. */
static inline int rc_frozen(rangecoder *rc)
{
    g_ora_symbols++;
    if (rc->poisoned) rc->failed = 1;
    return rc->failed;
}

static inline uint32_t rc_get_freq(rangecoder *rc, uint32_t tot)
{
    rc->range = rc->range / tot;                           /* >= 2^24 / (2^16 + step) on every path that gets here */
    if (rc->range == 0) { rc->failed = 1; return 0xFFFFFFFFu; }   /* (unreachable: a failed symbol freezes the coder) */
    uint64_t v = rc->code / rc->range;
    return v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
}

/* RangeCoder.hx:36-43 */
static inline void rc_decode(rangecoder *rc, uint32_t cum, uint32_t freq)
{
    rc->last_cum = cum; rc->last_freq = freq;
    rc->code -= (uint64_t)cum * rc->range;
    rc->range = rc->range * freq;
    while (rc->range < TOP) { rc_byte(rc); rc->range <<= 8; }
}

/* RangeCoder.hx:51-80 */
static int rc_decode_val(rangecoder *rc, uint32_t *cnt, int maxc, uint32_t step)
{
    if (rc_frozen(rc)) return 0;
    uint32_t totfr = cnt[maxc];
    rc->last_tot = totfr;
    uint32_t value = rc_get_freq(rc, totfr);
    int c = 0; uint32_t cumfr = 0, cnt_c = 0;
    while (c < maxc) {
        cnt_c = cnt[c];
        if (value >= cumfr + cnt_c) cumfr += cnt_c; else break;
        c++;
    }
    if (c == maxc) { rc->failed = 1; return maxc - 1; }      /* search ran off the table: not a valid stream */
    rc_decode(rc, cumfr, cnt_c);
    cnt[c] = cnt_c + step;
    totfr += step;
    if (totfr > BOT) {
        totfr = 0;
        for (int i = 0; i < maxc; i++) { uint32_t nc = (cnt[i] >> 1) + 1; cnt[i] = nc; totfr += nc; }
    }
    cnt[maxc] = totfr;
    return c;
}

/* RangeCoder.hx:82-130: 16 group sums at [off..off+15], total at [off+16], 256 counts at [off+17..] */
static int rc_decode_val_uni(rangecoder *rc, uint32_t *cnt, uint32_t step)
{
    if (rc_frozen(rc)) return 0;
    uint32_t totfr = cnt[16];
    rc->last_tot = totfr;
    uint32_t value = rc_get_freq(rc, totfr);
    int x = 0; uint32_t cumfr = 0, cnt_x = 0;
    while (x < 16) {
        cnt_x = cnt[x];
        if (value >= cumfr + cnt_x) cumfr += cnt_x; else break;
        x++;
    }
    int c = x * 16; uint32_t cnt_c = 0;
    while (c < 256) {
        cnt_c = cnt[c + 17];
        if (value >= cumfr + cnt_c) cumfr += cnt_c; else break;
        c++;
    }
    if (x == 16 || c == 256) { rc->failed = 1; return 255; }
    rc_decode(rc, cumfr, cnt_c);
    cnt[c + 17] = cnt_c + step;
    cnt[x] = cnt_x + step;
    totfr += step;
    if (totfr > BOT) {
        totfr = 0;
        for (int i = 17; i < 256 + 17; i++) { uint32_t nc = (cnt[i] >> 1) + 1; cnt[i] = nc; totfr += nc; }
        for (int i = 0; i < 16; i++) {
            uint32_t sum = 0;
            for (int j = 0; j < 16; j++) sum += cnt[(i << 4) + 17 + j];
            cnt[i] = sum;
        }
    }
    cnt[16] = totfr;
    return c;
}

/* ---- EntroCoderRC ---- */
enum { SC_STEP = 400, SC_NSTEP = 400, SC_BTSTEP = 10, SC_BTNSTEP = 20, SC_SXYSTEP = 100, SC_MSTEP = 100,
       SC_UNSTEP = 1000, SC_XXSTEP = 1, CNTABSZ = 273 };   /* EntroCoders.hx:43-51 */

typedef struct {
    entro base;
    rangecoder rc;
    uint32_t *cntab;                       /* 3 * 4096 * 273 */
    uint32_t ptypetab[6][7];
    uint32_t ntab[CC_NCXMAX][257];
    uint32_t xxtab[257], ntab2[257], bttab[6];
    uint32_t sxytab[4][17];
    uint32_t mvtab[2][SP_MSR_X * 2 + 1];
} entro_rc;

static void erc_destroy(entro *e) { entro_rc *r = (entro_rc *)e; free(r->cntab); free(r); }

/* EntroCoders.hx:74-79 */
static void erc_preinit(entro *e)
{
    entro_rc *r = (entro_rc *)e;
    for (int i = 0; i < 3 * CC_CXMAX; i++) r->cntab[(size_t)i * CNTABSZ + 16] = 0;
}

/* EntroCoders.hx:81-130 */
static void erc_renewI(entro *e)
{
    entro_rc *r = (entro_rc *)e;
    for (int i = 0; i < 3 * CC_CXMAX; i++) {
        uint32_t *p = r->cntab + (size_t)i * CNTABSZ;
        if (p[16] != 256) {
            for (int k = 0; k < 256; k++) p[k + 17] = 1;
            for (int k = 0; k < 16; k++) p[k] = 16;
            p[16] = 256;
        }
    }
    for (int n = 0; n < CC_NCXMAX; n++) { for (int i = 0; i < 256; i++) r->ntab[n][i] = 1; r->ntab[n][256] = 256; }
    for (int c = 0; c < 6; c++) { for (int i = 0; i < 6; i++) r->ptypetab[c][i] = 1; r->ptypetab[c][6] = 6; }
    for (int i = 0; i < 256; i++) { r->xxtab[i] = 1; r->ntab2[i] = 1; }
    r->xxtab[256] = 256; r->ntab2[256] = 256;
    for (int i = 0; i < 5; i++) r->bttab[i] = 1;
    r->bttab[5] = 5;
    for (int c = 0; c < 4; c++) { for (int i = 0; i < 16; i++) r->sxytab[c][i] = 1; r->sxytab[c][16] = 16; }
    for (int i = 0; i < SP_MSR_X * 2; i++) r->mvtab[0][i] = 1;
    r->mvtab[0][SP_MSR_X * 2] = SP_MSR_X * 2;
    for (int i = 0; i < SP_MSR_Y * 2; i++) r->mvtab[1][i] = 1;
    r->mvtab[1][SP_MSR_Y * 2] = SP_MSR_Y * 2;
}

static void erc_begin(entro *e, const uint8_t *src, int len, int pos0) { rc_begin(&((entro_rc *)e)->rc, src, len, pos0); }
#define ERC_T(kind, expr) do { int c_ = (expr); if (g_ora_sym_trace && !r->rc.failed) ora_sym_trace_put(kind, c_, (int)r->rc.last_freq, (int)r->rc.last_cum, (int)r->rc.last_tot); return c_; } while (0)
static int erc_clr(entro *e, int cxi) { entro_rc *r = (entro_rc *)e; ERC_T(0, rc_decode_val_uni(&r->rc, r->cntab + (size_t)cxi * CNTABSZ, SC_STEP)); }
static int erc_n(entro *e, int pt) { entro_rc *r = (entro_rc *)e; ERC_T(1, rc_decode_val(&r->rc, r->ntab[pt], 256, SC_NSTEP)); }
static int erc_p(entro *e, int pt) { entro_rc *r = (entro_rc *)e; ERC_T(2, rc_decode_val(&r->rc, r->ptypetab[pt], 6, SC_UNSTEP)); }
static int erc_x(entro *e) { entro_rc *r = (entro_rc *)e; ERC_T(3, rc_decode_val(&r->rc, r->xxtab, 256, SC_XXSTEP)); }
static int erc_bt(entro *e) { entro_rc *r = (entro_rc *)e; ERC_T(4, rc_decode_val(&r->rc, r->bttab, 5, SC_BTSTEP)); }
static int erc_bn(entro *e) { entro_rc *r = (entro_rc *)e; ERC_T(5, rc_decode_val(&r->rc, r->ntab2, 256, SC_BTNSTEP)); }
static int erc_sxy(entro *e, int n) { entro_rc *r = (entro_rc *)e; ERC_T(6, rc_decode_val(&r->rc, r->sxytab[n], 16, SC_SXYSTEP)); }
static int erc_mx(entro *e) { entro_rc *r = (entro_rc *)e; ERC_T(7, rc_decode_val(&r->rc, r->mvtab[0], SP_MSR_X * 2, SC_MSTEP)); }
static int erc_my(entro *e) { entro_rc *r = (entro_rc *)e; ERC_T(8, rc_decode_val(&r->rc, r->mvtab[1], SP_MSR_Y * 2, SC_MSTEP)); }
static int erc_canbool(entro *e) { (void)e; return 0; }
static int erc_bool(entro *e) { (void)e; return 0; }
static int erc_diff16(entro *e) { (void)e; return 1; }       /* EntroCoders.hx:72 */
static int erc_failed(entro *e) { return ((entro_rc *)e)->rc.failed; }
static void erc_fail(entro *e) { ((entro_rc *)e)->rc.failed = 1; }

entro *entro_rc_new(void)
{
    entro_rc *r = (entro_rc *)calloc(1, sizeof *r);
    r->cntab = (uint32_t *)calloc((size_t)3 * CC_CXMAX * CNTABSZ, sizeof(uint32_t));
    r->base.destroy = erc_destroy; r->base.preinit = erc_preinit; r->base.renewI = erc_renewI;
    r->base.decodeBegin = erc_begin; r->base.decodeClr = erc_clr; r->base.decodeN = erc_n; r->base.decodeP = erc_p;
    r->base.decodeX = erc_x; r->base.decodeBT = erc_bt; r->base.decodeBN = erc_bn; r->base.decodeSXY = erc_sxy;
    r->base.decodeMX = erc_mx; r->base.decodeMY = erc_my; r->base.canDecodeBool = erc_canbool; r->base.decodeBool = erc_bool;
    r->base.differentConstantsFor16bpp = erc_diff16; r->base.failed = erc_failed; r->base.fail = erc_fail;
    return &r->base;
}

/* Test hook for the model-level known-answer vector G5: a fresh colour context decodes
 * `value` to the symbol of the same number and then holds cnt[17+v]=401, cnt[v>>4]=416, cnt[16]=656. */
int ora_kat_rc_fresh_row(const uint8_t *src, int len, int cxi, uint32_t out[3])
{
    entro *e = entro_rc_new();
    entro_rc *r = (entro_rc *)e;
    e->preinit(e); e->renewI(e);
    e->decodeBegin(e, src, len, 1);
    int c = e->decodeClr(e, cxi);
    const uint32_t *row = r->cntab + (size_t)cxi * CNTABSZ;
    out[0] = row[17 + c]; out[1] = row[c >> 4]; out[2] = row[16];
    e->destroy(e);
    return c;
}
