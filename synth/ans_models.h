/*
 * ans_models.h -- CPU restatement of reference src/ANS.hx: the rANS state machine (Rans :5-49), the fixed-size
 * adaptive table (FixedSizeRansCtx :54-145) and the escalating colour-context kinds (SymbList/Cx1-3 :155-208,
 * SmallContext/Cx4/Cx5 :210-392, Cx6 :394-704, Cx7 :706-772, Context :785-860, Sorter :862-872).
 * Used by the synthetic rANS encoder (sp_ans_enc.c); the CPU checker compiles the same models for its decoder.
 * Not part of libjsplayer_cuda (the CUDA decoder csrc/sp_ans.cuh is an independent implementation).
 *
 * JavaScript typed-array semantics are explicit in the field types (Uint8Array -> uint8_t, Uint16Array -> uint16_t);
 * the reference's process-global statics (Context.rcv, SmallContext.totFr, Cx6.f0, Cx6._cnts/_freqs; ANS.hx:217,
 * 401-402,409,787) are per-call / per-decoder values here.
 */
#ifndef JSP_ANS_MODELS_H
#define JSP_ANS_MODELS_H
#include <stdint.h>
#include <string.h>

#define ANS_PROB_SCALE 4096
#define ANS_B 131072            /* Rans.B, ANS.hx:10 */
#define ANS_BYTE_L (1 << 23)    /* ANS.hx:33 */

typedef struct { int c, freq, cumFreq; } dec_receiver;   /* ANS.hx:149-153 */

/* ---- FixedSizeRansCtx, ANS.hx:54-145 ---- */
#define FX_STEP 16
#define FX_DSHIFT 7
#define FX_D (1 << FX_DSHIFT)
typedef struct {
    uint16_t freqs[512 * 2];   /* (freq, cumFreq) pairs */
    uint16_t cnts[512];
    int cntsum;
    uint8_t decTable[32];
    int NSym;
} fixed_ctx;

static inline void fx_set(fixed_ctx *t, int i, int fr, int cf) { t->freqs[i * 2] = (uint16_t)fr; t->freqs[i * 2 + 1] = (uint16_t)cf; }
static inline void fx_fill_dec(fixed_ctx *t, int i, int cf, int fr)
{
    int k0 = (cf + FX_D - 1) >> FX_DSHIFT, k1 = ((cf + fr - 1) >> FX_DSHIFT) + 1;
    for (int k = k0; k < k1; k++) if (k >= 0 && k < 32) t->decTable[k] = (uint8_t)i;
}
static inline void fx_renew(fixed_ctx *t)             /* :128-144 */
{
    int cf = 0, fr = ANS_PROB_SCALE / t->NSym, c0 = fr - (fr >> 1);
    t->cntsum = c0 * t->NSym;
    for (int i = 0; i < t->NSym; i++) { fx_set(t, i, fr, cf); t->cnts[i] = (uint16_t)c0; fx_fill_dec(t, i, cf, fr); cf += fr; }
}
static inline void fx_init(fixed_ctx *t, int nsym) { memset(t, 0, sizeof *t); t->NSym = nsym; }
static inline void fx_incr(fixed_ctx *t, int c)       /* :85-103 */
{
    t->cnts[c] = (uint16_t)(t->cnts[c] + FX_STEP); t->cntsum += FX_STEP;
    if (t->cntsum + FX_STEP > ANS_PROB_SCALE) {
        t->cntsum = 0; int cf = 0;
        for (int j = 0; j < t->NSym; j++) {
            int fr = t->cnts[j];
            fx_set(t, j, fr, cf);
            fx_fill_dec(t, j, cf, fr);
            cf += fr;
            t->cnts[j] = (uint16_t)(t->cnts[j] - (fr >> 1));
            t->cntsum += t->cnts[j];
        }
    }
}
static inline void fx_decode(fixed_ctx *t, int someFreq, dec_receiver *r)   /* :105-126 */
{
    int c0 = t->decTable[someFreq >> FX_DSHIFT];
    for (int j = c0; j < t->NSym - 1; j++)
        if (t->freqs[(j + 1) * 2 + 1] > someFreq) {
            r->freq = t->freqs[j * 2]; r->cumFreq = t->freqs[j * 2 + 1]; r->c = j;
            fx_incr(t, j);
            return;
        }
    int l = t->NSym - 1;
    r->freq = t->freqs[l * 2]; r->cumFreq = t->freqs[l * 2 + 1]; r->c = l;
    fx_incr(t, l);
}

/* ---- colour contexts ---- */
enum { CXK_NONE = 0, CXK_1, CXK_2, CXK_3, CXK_4, CXK_5, CXK_6, CXK_7 };

typedef struct {
    int kind;
    /* Cx1/Cx2/Cx3 (SymbList :155-208): raw symbols seen so far */
    uint8_t symb[256]; int d;
    /* SmallContext (Cx4: S=4, Cx5: S=16), :210-392 */
    int maxpos, S, cntsum5;
    uint8_t sc_symbols[16]; uint16_t sc_freqs[16];
    /* Cx6, :394-704 */
    int S6, fshift;
    uint8_t c6_symbols[64]; uint16_t c6_freqs[128]; uint16_t c6_cnts[65];
    /* Cx7 */
    fixed_ctx *c7;      /* allocated on demand */
} color_ctx;

void cctx_renew(color_ctx *x);                                   /* Context.renew :793 */
void cctx_free(color_ctx *x);
/* Context.decode :795-810 -- returns 0 for the raw kinds (None,1,2,3), else 1 with r filled and stats updated */
int  cctx_decode(color_ctx *x, int someFreq, dec_receiver *r, int f0);
void cctx_update(color_ctx *x, int c, int f0);                   /* Context.update :812-829 */
/* encoder side: the interval the decoder will compute for symbol c (no state change); 0 if raw kind;
 * -1 if the symbol cannot be coded (interval outside the 12-bit code space) */
int  cctx_interval(const color_ctx *x, int c, int *freq, int *cum);

#endif
