/*
 * ans_oracle.c -- CPU restatement of EntroCoderANS (reference src/EntroCoders.hx:182-313) with the byte-wise rANS
 * decoder (src/ANS.hx:5-49).  The adaptive models of ANS.hx (:54-872) are restated in
 * synth/ans_models.{h,c} -- next to the synthetic rANS encoder, which has to run the very same models
 *  -- and are compiled into liboracle.so from there (oracle/Makefile).  The dependency points
 * from the oracle to the encoder's models, never from the product to the oracle.
 * TEST INFRASTRUCTURE ONLY.
 *
 * JavaScript semantics made explicit: the rANS state is an int32 after every `<<`/`|`/`&`/`>>` (ANS.hx:25-29,
 * 39-42); an out-of-bounds byte is `undefined`, which `|` turns into 0.  Defined behaviour (not in the reference):
 * asking for a symbol after the reader has run past the end, or a renormalisation that cannot terminate, is
 * reported through failed() -- no valid stream does either.
 */
#include "../synth/ans_models.h"
#include "sp_entro.h"
#include <stdlib.h>

/* ---------------------------------------------------------------- rANS + EntroCoderANS ---- */
typedef struct {
    int64_t r; int pos, len; const uint8_t *data;
    int overrun, failed;
} rans_t;

static inline int rans_byte(rans_t *s)
{
    int b = 0;
    if (s->pos >= 0 && s->pos < s->len) b = s->data[s->pos]; else s->overrun = 1;
    s->pos++;
    return b;
}
static void rans_init(rans_t *s, int i)                    /* reinitImpl, ANS.hx:22-31 */
{
    s->pos = i;
    uint32_t x = (uint32_t)rans_byte(s);
    x |= (uint32_t)rans_byte(s) << 8; x |= (uint32_t)rans_byte(s) << 16; x |= (uint32_t)rans_byte(s) << 24;
    s->r = (int32_t)x;
}
static inline int rans_get(rans_t *s) { if (s->overrun) s->failed = 1; return (int)((uint32_t)s->r & 4095u); }   /* :35 */
static void rans_advance(rans_t *s, int start, int freq)   /* :37-44 */
{
    const int32_t r32 = (int32_t)(uint32_t)s->r;
    int64_t x = (int64_t)freq * (r32 >> 12) + (r32 & 4095) - start;
    int guard = 0;
    while (x < ANS_BYTE_L) {
        if (s->overrun || ++guard > 8) { s->failed = 1; break; }
        x = (int32_t)(((uint32_t)x << 8) | (uint32_t)rans_byte(s));
    }
    s->r = x;
}

typedef struct {
    entro base;
    rans_t rans;
    int nDec, f0;
    color_ctx *cntab;            /* 3 * 4096 */
    fixed_ctx ptypetab[6], ntab[CC_NCXMAX], xxtab, ntab2, bttab, sxytab[4], mvtab[2];
} entro_ans;

static void ea_destroy(entro *e)
{
    entro_ans *a = (entro_ans *)e;
    for (int i = 0; i < 3 * CC_CXMAX; i++) cctx_free(&a->cntab[i]);
    free(a->cntab); free(a);
}
static void ea_preinit(entro *e) { (void)e; }
static void ea_renewI(entro *e)                            /* EntroCoders.hx:216-227 */
{
    entro_ans *a = (entro_ans *)e;
    for (int i = 0; i < 3 * CC_CXMAX; i++) cctx_renew(&a->cntab[i]);
    for (int i = 0; i < CC_NCXMAX; i++) fx_renew(&a->ntab[i]);
    for (int i = 0; i < 6; i++) fx_renew(&a->ptypetab[i]);
    fx_renew(&a->xxtab); fx_renew(&a->ntab2); fx_renew(&a->bttab);
    for (int i = 0; i < 4; i++) fx_renew(&a->sxytab[i]);
    for (int i = 0; i < 2; i++) fx_renew(&a->mvtab[i]);
}
static void ea_begin(entro *e, const uint8_t *src, int len, int pos0)   /* :229-233 */
{
    entro_ans *a = (entro_ans *)e;
    a->rans.data = src; a->rans.len = len; a->rans.overrun = 0;
    a->rans.failed = 0;                                   /* the failure report is per frame */
    rans_init(&a->rans, pos0);
    a->nDec = 0;
}
extern __thread unsigned long long g_ora_symbols;
extern int32_t *g_ora_sym_trace; void ora_sym_trace_put(int kind, int c, int freq, int cum, int tot);
static inline void ea_count(entro_ans *a)
{
    g_ora_symbols++;
    a->nDec++;
    if (a->nDec == ANS_B) { rans_init(&a->rans, a->rans.pos); a->nDec = 0; }
}
static int ea_clr(entro *e, int cxi)                       /* :235-255 */
{
    entro_ans *a = (entro_ans *)e;
    color_ctx *dcx = &a->cntab[cxi];
    dec_receiver rcv; int c;
    if (cctx_decode(dcx, rans_get(&a->rans), &rcv, a->f0)) {
        c = rcv.c;
        rans_advance(&a->rans, rcv.cumFreq, rcv.freq);
        /* defined behaviour: an escape interval past symbol 255 (impossible on a valid stream; the reference would
         * index cntab[] out of range on the next symbol and throw) is a failure and the symbol wraps to a byte */
        if (c > 255) { a->rans.failed = 1; c &= 255; }
        if (g_ora_sym_trace) ora_sym_trace_put(0, c, rcv.freq, rcv.cumFreq, 4096);
    } else {
        c = rans_byte(&a->rans);                           /* Rans.raw, ANS.hx:46-48 */
        cctx_update(dcx, c, a->f0);
        if (g_ora_sym_trace) ora_sym_trace_put(0, c, 0, 0, 0);
    }
    ea_count(a);
    return c;
}
static int ea_bool(entro *e)                               /* :259-269 */
{
    entro_ans *a = (entro_ans *)e;
    const int f = rans_get(&a->rans);
    const int flag = f >= (ANS_PROB_SCALE >> 1);
    rans_advance(&a->rans, flag ? ANS_PROB_SCALE >> 1 : 0, ANS_PROB_SCALE >> 1);
    if (g_ora_sym_trace) ora_sym_trace_put(9, flag, ANS_PROB_SCALE >> 1, flag ? ANS_PROB_SCALE >> 1 : 0, 4096);
    ea_count(a);
    return flag;
}
static int ea_f(entro_ans *a, fixed_ctx *t, int kind)      /* decodeF, :271-280 */
{
    dec_receiver rcv;
    fx_decode(t, rans_get(&a->rans), &rcv);
    rans_advance(&a->rans, rcv.cumFreq, rcv.freq);
    if (g_ora_sym_trace) ora_sym_trace_put(kind, rcv.c, rcv.freq, rcv.cumFreq, 4096);
    ea_count(a);
    return rcv.c;
}
static int ea_n(entro *e, int pt) { entro_ans *a = (entro_ans *)e; return ea_f(a, &a->ntab[pt], 1); }
static int ea_p(entro *e, int pt) { entro_ans *a = (entro_ans *)e; return ea_f(a, &a->ptypetab[pt], 2); }
static int ea_x(entro *e) { entro_ans *a = (entro_ans *)e; return ea_f(a, &a->xxtab, 3); }
static int ea_bt(entro *e) { entro_ans *a = (entro_ans *)e; return ea_f(a, &a->bttab, 4); }
static int ea_bn(entro *e) { entro_ans *a = (entro_ans *)e; return ea_f(a, &a->ntab2, 5); }
static int ea_sxy(entro *e, int n) { entro_ans *a = (entro_ans *)e; return ea_f(a, &a->sxytab[n], 6); }
static int ea_mx(entro *e) { entro_ans *a = (entro_ans *)e; return ea_f(a, &a->mvtab[0], 7); }
static int ea_my(entro *e) { entro_ans *a = (entro_ans *)e; return ea_f(a, &a->mvtab[1], 8); }
static int ea_canbool(entro *e) { (void)e; return 1; }
static int ea_diff16(entro *e) { (void)e; return 0; }      /* EntroCoders.hx:214 */
static int ea_failed(entro *e) { return ((entro_ans *)e)->rans.failed; }
static void ea_fail(entro *e) { ((entro_ans *)e)->rans.failed = 1; }

entro *entro_ans_new(int f0val)
{
    entro_ans *a = (entro_ans *)calloc(1, sizeof *a);
    a->cntab = (color_ctx *)calloc((size_t)3 * CC_CXMAX, sizeof(color_ctx));
    a->f0 = f0val;
    for (int i = 0; i < CC_NCXMAX; i++) fx_init(&a->ntab[i], 256);
    for (int i = 0; i < 6; i++) fx_init(&a->ptypetab[i], 6);
    fx_init(&a->xxtab, 256); fx_init(&a->ntab2, 256); fx_init(&a->bttab, 5);
    for (int i = 0; i < 4; i++) fx_init(&a->sxytab[i], 16);
    for (int i = 0; i < 2; i++) fx_init(&a->mvtab[i], 512);
    a->base.destroy = ea_destroy; a->base.preinit = ea_preinit; a->base.renewI = ea_renewI; a->base.decodeBegin = ea_begin;
    a->base.decodeClr = ea_clr; a->base.decodeN = ea_n; a->base.decodeP = ea_p; a->base.decodeX = ea_x; a->base.decodeBT = ea_bt;
    a->base.decodeBN = ea_bn; a->base.decodeSXY = ea_sxy; a->base.decodeMX = ea_mx; a->base.decodeMY = ea_my;
    a->base.canDecodeBool = ea_canbool; a->base.decodeBool = ea_bool; a->base.differentConstantsFor16bpp = ea_diff16;
    a->base.failed = ea_failed; a->base.fail = ea_fail;
    return &a->base;
}

/* Test hook for the model-level known-answer vector G6. out[0..5] =
 * kind after 1st symbol, kind after repeating it, d, freqs[0], and for FixedSizeRansCtx(256).renew():
 * freq/cum of symbol 3 packed, cnt, cntsum, decTable[5]. */
void ora_kat_ans(int sym, int out[10])
{
    color_ctx x; memset(&x, 0, sizeof x);
    cctx_renew(&x);
    dec_receiver r;
    out[0] = cctx_decode(&x, 0, &r, 32);          /* fresh context: raw */
    cctx_update(&x, sym, 32);
    out[1] = x.kind;
    cctx_update(&x, sym, 32);
    out[2] = x.kind; out[3] = x.d; out[4] = x.sc_freqs[0];
    fixed_ctx t; fx_init(&t, 256); fx_renew(&t);
    out[5] = t.freqs[3 * 2]; out[6] = t.freqs[3 * 2 + 1]; out[7] = t.cnts[3]; out[8] = t.cntsum; out[9] = t.decTable[5];
    cctx_free(&x);
}
