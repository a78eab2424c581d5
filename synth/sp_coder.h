/* sp_coder.h -- encoder-side entropy-coder interface of the synthetic ScreenPressor encoder.
 * One implementation per stream version: sp_rc_enc.c (v2, range coder) and sp_ans_enc.c (v3/v4, rANS).
 * The symbol vocabulary is the decoder's (reference src/EntroCoders.hx:8-24). */
#ifndef JSP_SP_CODER_H
#define JSP_SP_CODER_H
#include <stddef.h>
#include <stdint.h>

typedef struct sp_coder sp_coder;
struct sp_coder {
    void (*destroy)(sp_coder *);
    void (*renew_i)(sp_coder *);                 /* model reset at a coded I frame */
    void (*begin)(sp_coder *);                   /* start of a frame payload */
    void (*clr)(sp_coder *, int cxi, int sym);   /* colour channel symbol in context cxi (0..12287) */
    void (*n)(sp_coder *, int ptype, int sym);
    void (*p)(sp_coder *, int prev_ptype, int sym);
    void (*x)(sp_coder *, int sym);
    void (*bt)(sp_coder *, int sym);
    void (*bn)(sp_coder *, int sym);
    void (*sxy)(sp_coder *, int k, int sym);
    void (*mx)(sp_coder *, int sym);
    void (*my)(sp_coder *, int sym);
    int  (*can_bool)(sp_coder *);
    void (*boolean)(sp_coder *, int flag);
    /* finishes the payload; returns its size (bytes written to out) or 0 if cap is too small / the frame
     * cannot be represented (rANS: an interval outside the 12-bit code space) */
    size_t (*finish)(sp_coder *, uint8_t *out, size_t cap);
};

sp_coder *sp_rc_coder_new(void);
sp_coder *sp_ans_coder_new(int f0);

#endif
