/* placeholder until the ScreenPressor restatement lands (replaced below in history) */
#include "oracle_internal.h"
#include <stdlib.h>
struct sp_dec { int X, Y, bpp; };
sp_dec *sp_new(int w, int h, int bpp) { sp_dec *s = calloc(1, sizeof *s); s->X = w; s->Y = h; s->bpp = bpp; return s; }
void sp_free(sp_dec *s) { free(s); }
void sp_preinit(sp_dec *s, int n) { (void)s; (void)n; }
int sp_is_key(const uint8_t *src, int len) { (void)src; (void)len; return 0; }
const int32_t *sp_prev(sp_dec *s) { (void)s; return 0; }
int sp_decompress_i(sp_dec *s, const uint8_t *src, int len, int32_t *dst) { (void)s;(void)src;(void)len;(void)dst; return ORA_ERROR_OCCURED; }
int sp_decompress_p(sp_dec *s, const uint8_t *src, int len, int32_t *dst, const int32_t **p, int *sig) { (void)s;(void)src;(void)len;(void)dst; *p = 0; *sig = 0; return ORA_ERROR_OCCURED; }
void sp_stop(sp_dec *s) { (void)s; }
