/*
 * sp_ans_enc.c -- rANS ENCODER for synthetic ScreenPressor v3/v4 streams.  It drives the
 * decoder-side models of ans_models.{h,c} symbol by symbol, so the intervals it codes are the ones a decoder that
 * follows reference src/ANS.hx computes.
 *
 * Two passes per frame: forward over the symbols collecting (cumFreq, freq) or a raw byte per symbol while the
 * models adapt exactly as the decoder's will; then, per block of Rans.B symbols (the decoder re-reads its state
 * every B symbols, EntroCoders.hx:249-253), a reverse pass of byte-wise rANS encoding whose output is reversed.
 */
#include "ans_models.h"
#include "sp_coder.h"
#include <stdlib.h>

typedef struct { uint16_t cum, freq; } ans_sym;      /* freq == 0: raw byte in `cum` */

typedef struct {
    sp_coder base;
    int f0, bad;
    color_ctx *cntab;
    fixed_ctx ptypetab[6], ntab[6], xxtab, ntab2, bttab, sxytab[4], mvtab[2];
    ans_sym *syms; size_t n, cap;
} ans_coder;

static void push(ans_coder *a, int cum, int freq)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 65536; a->syms = (ans_sym *)realloc(a->syms, a->cap * sizeof(ans_sym)); }
    a->syms[a->n].cum = (uint16_t)cum; a->syms[a->n].freq = (uint16_t)freq; a->n++;
}
static void ac_destroy(sp_coder *b)
{
    ans_coder *a = (ans_coder *)b;
    for (int i = 0; i < 12288; i++) cctx_free(&a->cntab[i]);
    free(a->cntab); free(a->syms); free(a);
}
static void ac_renew(sp_coder *b)
{
    ans_coder *a = (ans_coder *)b;
    for (int i = 0; i < 12288; i++) cctx_renew(&a->cntab[i]);
    for (int i = 0; i < 6; i++) { fx_renew(&a->ntab[i]); fx_renew(&a->ptypetab[i]); }
    fx_renew(&a->xxtab); fx_renew(&a->ntab2); fx_renew(&a->bttab);
    for (int i = 0; i < 4; i++) fx_renew(&a->sxytab[i]);
    fx_renew(&a->mvtab[0]); fx_renew(&a->mvtab[1]);
}
static void ac_begin(sp_coder *b) { ans_coder *a = (ans_coder *)b; a->n = 0; a->bad = 0; }
static void ac_clr(sp_coder *b, int cxi, int sym)
{
    ans_coder *a = (ans_coder *)b;
    color_ctx *x = &a->cntab[cxi];
    int fr = 0, cum = 0;
    const int k = cctx_interval(x, sym, &fr, &cum);
    if (k == 0) { push(a, sym, 0); cctx_update(x, sym, a->f0); return; }
    if (k < 0) { a->bad = 1; return; }
    dec_receiver r;
    cctx_decode(x, cum, &r, a->f0);                   /* the decoder's own update; must agree on the interval */
    if (r.c != sym || r.freq != fr || r.cumFreq != cum) a->bad = 1;
    push(a, cum, fr);
}
static void ac_fixed(ans_coder *a, fixed_ctx *t, int sym)
{
    const int fr = t->freqs[sym * 2], cum = t->freqs[sym * 2 + 1];
    if (fr == 0 || cum + fr > ANS_PROB_SCALE) { a->bad = 1; return; }
    dec_receiver r;
    fx_decode(t, cum, &r);
    if (r.c != sym) a->bad = 1;
    push(a, cum, fr);
}
static void ac_n(sp_coder *b, int pt, int s) { ans_coder *a = (ans_coder *)b; ac_fixed(a, &a->ntab[pt], s); }
static void ac_p(sp_coder *b, int pt, int s) { ans_coder *a = (ans_coder *)b; ac_fixed(a, &a->ptypetab[pt], s); }
static void ac_x(sp_coder *b, int s) { ans_coder *a = (ans_coder *)b; ac_fixed(a, &a->xxtab, s); }
static void ac_bt(sp_coder *b, int s) { ans_coder *a = (ans_coder *)b; ac_fixed(a, &a->bttab, s); }
static void ac_bn(sp_coder *b, int s) { ans_coder *a = (ans_coder *)b; ac_fixed(a, &a->ntab2, s); }
static void ac_sxy(sp_coder *b, int k, int s) { ans_coder *a = (ans_coder *)b; ac_fixed(a, &a->sxytab[k], s); }
static void ac_mx(sp_coder *b, int s) { ans_coder *a = (ans_coder *)b; ac_fixed(a, &a->mvtab[0], s); }
static void ac_my(sp_coder *b, int s) { ans_coder *a = (ans_coder *)b; ac_fixed(a, &a->mvtab[1], s); }
static int ac_can_bool(sp_coder *b) { (void)b; return 1; }
static void ac_boolean(sp_coder *b, int flag) { push((ans_coder *)b, flag ? 2048 : 0, 2048); }

static size_t ac_finish(sp_coder *b, uint8_t *out, size_t cap)
{
    ans_coder *a = (ans_coder *)b;
    if (a->bad) return 0;
    size_t o = 0;
    uint8_t *tmp = (uint8_t *)malloc(a->n * 3 + 64);
    for (size_t b0 = 0; b0 < a->n; b0 += ANS_B) {
        const size_t b1 = b0 + ANS_B < a->n ? b0 + ANS_B : a->n;
        size_t t = 0;
        uint32_t x = ANS_BYTE_L;
        for (size_t i = b1; i-- > b0;) {
            const ans_sym s = a->syms[i];
            if (s.freq == 0) { tmp[t++] = (uint8_t)s.cum; continue; }
            const uint32_t x_max = ((ANS_BYTE_L >> 12) << 8) * s.freq;
            while (x >= x_max) { tmp[t++] = (uint8_t)(x & 0xFF); x >>= 8; }
            x = ((x / s.freq) << 12) + (x % s.freq) + s.cum;
        }
        tmp[t++] = (uint8_t)(x >> 24); tmp[t++] = (uint8_t)(x >> 16); tmp[t++] = (uint8_t)(x >> 8); tmp[t++] = (uint8_t)x;
        if (o + t > cap) { free(tmp); return 0; }
        for (size_t i = 0; i < t; i++) out[o + i] = tmp[t - 1 - i];
        o += t;
    }
    free(tmp);
    return o;
}

sp_coder *sp_ans_coder_new(int f0)
{
    ans_coder *a = (ans_coder *)calloc(1, sizeof *a);
    a->f0 = f0;
    a->cntab = (color_ctx *)calloc(12288, sizeof(color_ctx));
    for (int i = 0; i < 6; i++) { fx_init(&a->ntab[i], 256); fx_init(&a->ptypetab[i], 6); }
    fx_init(&a->xxtab, 256); fx_init(&a->ntab2, 256); fx_init(&a->bttab, 5);
    for (int i = 0; i < 4; i++) fx_init(&a->sxytab[i], 16);
    fx_init(&a->mvtab[0], 512); fx_init(&a->mvtab[1], 512);
    a->base.destroy = ac_destroy; a->base.renew_i = ac_renew; a->base.begin = ac_begin; a->base.clr = ac_clr;
    a->base.n = ac_n; a->base.p = ac_p; a->base.x = ac_x; a->base.bt = ac_bt; a->base.bn = ac_bn; a->base.sxy = ac_sxy;
    a->base.mx = ac_mx; a->base.my = ac_my; a->base.can_bool = ac_can_bool; a->base.boolean = ac_boolean; a->base.finish = ac_finish;
    ac_renew(&a->base);
    return &a->base;
}
