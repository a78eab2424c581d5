/* oracle_internal.h -- private glue between the oracle's translation units.
 * TEST INFRASTRUCTURE ONLY (see oracle.h). */
#ifndef JSP_ORACLE_INTERNAL_H
#define JSP_ORACLE_INTERNAL_H
#include "oracle.h"
#include <stddef.h>

typedef struct msv1_dec msv1_dec;
typedef struct sp_dec sp_dec;

struct ora_dec {
    int codec;
    int X, Y, bpp;
    msv1_dec *msv1;
    sp_dec *sp;
};

/* msvideo1_oracle.c */
msv1_dec *msv1_new(int is8, int w, int h, const uint8_t *palette, int palette_bytes);
void msv1_free(msv1_dec *m);
void msv1_preinit(msv1_dec *m, int insignificant_lines);
int  msv1_is_key(msv1_dec *m, const uint8_t *src, int len);
const int32_t *msv1_prev(msv1_dec *m);
void msv1_decompress_p(msv1_dec *m, const uint8_t *src, int len, int32_t *dst,
                       const int32_t **data_pnt, int *signif);

/* screenpressor_oracle.c */
sp_dec *sp_new(int w, int h, int bpp);
void sp_free(sp_dec *s);
void sp_preinit(sp_dec *s, int insignificant_lines);
int  sp_is_key(const uint8_t *src, int len);
const int32_t *sp_prev(sp_dec *s);
int  sp_decompress_i(sp_dec *s, const uint8_t *src, int len, int32_t *dst);
int  sp_decompress_p(sp_dec *s, const uint8_t *src, int len, int32_t *dst,
                     const int32_t **data_pnt, int *signif);
void sp_stop(sp_dec *s);

#endif
