/*
 * sp_rc_enc.c -- range ENCODER + adaptive frequency models for synthetic ScreenPressor v2 streams.
 *
 * The reference only has the decoder (src/RangeCoder.hx); its renormalisation (`while range < 2^24`,
 * :40-42) and the byte it skips at DecodeBegin (:29-33) pair with a carry-propagating encoder of the
 * LZMA "shiftLow" kind.  Model updates replay the decoder's (RangeCoder.hx:68-78
 * for plain tables, :110-128 for the 16x16 colour tables; steps and sizes EntroCoders.hx:43-69) so both
 * sides stay in lock-step.  The colour tables keep only the 256 counts + total: the decoder's 16 group
 * sums are derived data.
 */
#include "sp_coder.h"
#include <stdlib.h>
#include <string.h>

#define BOT 0x10000u

typedef struct {
    uint64_t low; uint32_t range; uint8_t cache; uint64_t cache_size;
    uint8_t *buf; size_t n, cap;
} renc;

static void renc_put(renc *e, uint8_t b)
{
    if (e->n == e->cap) { e->cap = e->cap ? e->cap * 2 : 65536; e->buf = (uint8_t *)realloc(e->buf, e->cap); }
    e->buf[e->n++] = b;
}
static void renc_shift_low(renc *e)
{
    if ((uint32_t)e->low < 0xFF000000u || (e->low >> 32) != 0) {
        uint8_t t = e->cache;
        do { renc_put(e, (uint8_t)(t + (uint8_t)(e->low >> 32))); t = 0xFF; } while (--e->cache_size != 0);
        e->cache = (uint8_t)(e->low >> 24);
    }
    e->cache_size++;
    e->low = (e->low & 0x00FFFFFFu) << 8;
}
static void renc_begin(renc *e) { e->low = 0; e->range = 0xFFFFFFFFu; e->cache = 0; e->cache_size = 1; e->n = 0; }
static void renc_encode(renc *e, uint32_t cum, uint32_t freq, uint32_t tot)
{
    uint32_t r = e->range / tot;
    e->low += (uint64_t)r * cum;
    e->range = r * freq;
    while (e->range < 0x01000000u) { e->range <<= 8; renc_shift_low(e); }
}

typedef struct {
    sp_coder base;
    renc rc;
    uint32_t *clr;                 /* 12288 rows x 257 (256 counts + total) */
    uint32_t *touched; size_t n_touched;
    uint8_t *dirty;
    uint32_t ptypetab[6][7], ntab[6][257], xxtab[257], ntab2[257], bttab[6], sxytab[4][17], mvtab[2][513];
} rc_coder;

static void table_reset(uint32_t *t, int n) { for (int i = 0; i < n; i++) t[i] = 1; t[n] = (uint32_t)n; }

static void enc_val(rc_coder *c, uint32_t *cnt, int maxc, uint32_t step, int sym)
{
    uint32_t cum = 0;
    for (int i = 0; i < sym; i++) cum += cnt[i];
    uint32_t tot = cnt[maxc];
    renc_encode(&c->rc, cum, cnt[sym], tot);
    cnt[sym] += step; tot += step;
    if (tot > BOT) { tot = 0; for (int i = 0; i < maxc; i++) { cnt[i] = (cnt[i] >> 1) + 1; tot += cnt[i]; } }
    cnt[maxc] = tot;
}

static void rc_destroy(sp_coder *b) { rc_coder *c = (rc_coder *)b; free(c->rc.buf); free(c->clr); free(c->touched); free(c->dirty); free(c); }
static void rc_renew(sp_coder *b)
{
    rc_coder *c = (rc_coder *)b;
    for (size_t k = 0; k < c->n_touched; k++) {
        uint32_t *row = c->clr + (size_t)c->touched[k] * 257;
        table_reset(row, 256); c->dirty[c->touched[k]] = 0;
    }
    c->n_touched = 0;
    for (int i = 0; i < 6; i++) { table_reset(c->ntab[i], 256); table_reset(c->ptypetab[i], 6); }
    table_reset(c->xxtab, 256); table_reset(c->ntab2, 256); table_reset(c->bttab, 5);
    for (int i = 0; i < 4; i++) table_reset(c->sxytab[i], 16);
    table_reset(c->mvtab[0], 512); table_reset(c->mvtab[1], 512);
}
static void rc_begin(sp_coder *b) { renc_begin(&((rc_coder *)b)->rc); }
static void rc_clr(sp_coder *b, int cxi, int sym)
{
    rc_coder *c = (rc_coder *)b;
    if (!c->dirty[cxi]) { c->dirty[cxi] = 1; c->touched[c->n_touched++] = (uint32_t)cxi; }
    enc_val(c, c->clr + (size_t)cxi * 257, 256, 400, sym);
}
static void rc_n(sp_coder *b, int pt, int s) { rc_coder *c = (rc_coder *)b; enc_val(c, c->ntab[pt], 256, 400, s); }
static void rc_p(sp_coder *b, int pt, int s) { rc_coder *c = (rc_coder *)b; enc_val(c, c->ptypetab[pt], 6, 1000, s); }
static void rc_x(sp_coder *b, int s) { rc_coder *c = (rc_coder *)b; enc_val(c, c->xxtab, 256, 1, s); }
static void rc_bt(sp_coder *b, int s) { rc_coder *c = (rc_coder *)b; enc_val(c, c->bttab, 5, 10, s); }
static void rc_bn(sp_coder *b, int s) { rc_coder *c = (rc_coder *)b; enc_val(c, c->ntab2, 256, 20, s); }
static void rc_sxy(sp_coder *b, int k, int s) { rc_coder *c = (rc_coder *)b; enc_val(c, c->sxytab[k], 16, 100, s); }
static void rc_mx(sp_coder *b, int s) { rc_coder *c = (rc_coder *)b; enc_val(c, c->mvtab[0], 512, 100, s); }
static void rc_my(sp_coder *b, int s) { rc_coder *c = (rc_coder *)b; enc_val(c, c->mvtab[1], 512, 100, s); }
static int rc_can_bool(sp_coder *b) { (void)b; return 0; }
static void rc_boolean(sp_coder *b, int f) { (void)b; (void)f; }
static size_t rc_finish(sp_coder *b, uint8_t *out, size_t cap)
{
    rc_coder *c = (rc_coder *)b;
    for (int i = 0; i < 5; i++) renc_shift_low(&c->rc);
    if (c->rc.n > cap) return 0;
    memcpy(out, c->rc.buf, c->rc.n);
    return c->rc.n;
}

sp_coder *sp_rc_coder_new(void)
{
    rc_coder *c = (rc_coder *)calloc(1, sizeof *c);
    c->clr = (uint32_t *)malloc((size_t)12288 * 257 * 4);
    c->touched = (uint32_t *)malloc(12288 * 4);
    c->dirty = (uint8_t *)calloc(12288, 1);
    for (int i = 0; i < 12288; i++) table_reset(c->clr + (size_t)i * 257, 256);
    c->base.destroy = rc_destroy; c->base.renew_i = rc_renew; c->base.begin = rc_begin; c->base.clr = rc_clr;
    c->base.n = rc_n; c->base.p = rc_p; c->base.x = rc_x; c->base.bt = rc_bt; c->base.bn = rc_bn; c->base.sxy = rc_sxy;
    c->base.mx = rc_mx; c->base.my = rc_my; c->base.can_bool = rc_can_bool; c->base.boolean = rc_boolean; c->base.finish = rc_finish;
    rc_renew(&c->base);
    return &c->base;
}
