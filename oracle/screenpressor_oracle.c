/*
 * screenpressor_oracle.c -- CPU restatement of reference src/ScreenPressor.hx (whole file).
 * TEST INFRASTRUCTURE ONLY.  Entropy coders: rangecoder_oracle.c (v2), ans_oracle.c (v3/v4).
 *
 * Parity pinning: no independent ScreenPressor decoder exists in this image and the reference ships no
 * vectors, so this restatement is "parity unpinned" beyond model-level KATs and encoder round trips
 * (tests/test_oracle_sp.py).
 *
 * Defined behaviour where the reference has none, identical in the CUDA path:
 *  - the reference's frame buffers are 4x too large (Manager.hx:114-118) and a final run may spill past
 *    X*Y; here pictures are exactly X*Y and writes past the end are dropped;
 *  - out-of-frame reads (negative indices, motion vectors leaving the picture) yield 0
 *    (JavaScript `undefined` stored into an Int32Array); predictor 4 adds BYTES of three neighbours, and there one
 *    `undefined` operand makes the whole sum NaN -> the pixel is 0 (not "the missing neighbour counts as 0");
 *  - a P frame starts from a copy of the previous picture (the reference copies unchanged blocks one by
 *    one, ScreenPressor.hx:468-474; every pixel of a valid frame is written either way);
 *  - a flat I frame before any coded I frame (ec == null, ScreenPressor.hx:112-114) and a stream whose
 *    entropy coder fails (see rangecoder_oracle.c) return error_occured.
 */
#include "oracle_internal.h"
#include "sp_entro.h"
#include <stdlib.h>
#include <string.h>

struct sp_dec {
    int X, Y, bpp;
    int cx, cx1;
    entro *ec;
    int SC_CXSHIFT;
    const int32_t *prevFrame;
    int nbx, nby;
    int32_t *bts;
    int insignificant_blocks;
    int decodedI;
    int last_one_was_flat;      /* Null<Int>: -1 = null */
    int decodingBools;
    int ctx_fail;               /* a colour context index left its channel's 4096 contexts (see ctx_index) */
    long budget;                /* run-loop iterations left in this frame (see SP_RUN_BUDGET) */
};

sp_dec *sp_new(int w, int h, int bpp)
{
    sp_dec *s = (sp_dec *)calloc(1, sizeof *s);
    s->X = w; s->Y = h; s->bpp = bpp;
    s->SC_CXSHIFT = bpp == 16 ? 0 : 2;                    /* ScreenPressor.hx:59 */
    s->nbx = (w + 15) / 16; s->nby = (h + 15) / 16;       /* :60-61 */
    s->bts = (int32_t *)calloc((size_t)s->nbx * s->nby + 1, sizeof(int32_t));
    s->last_one_was_flat = -1;
    return s;
}

void sp_free(sp_dec *s) { if (!s) return; if (s->ec) s->ec->destroy(s->ec); free(s->bts); free(s); }
void sp_preinit(sp_dec *s, int lines) { s->insignificant_blocks = s->nbx * ((lines + 15) / 16); }   /* :86-89 */
const int32_t *sp_prev(sp_dec *s) { return s->prevFrame; }
void sp_stop(sp_dec *s) { if (s->ec) { s->ec->destroy(s->ec); s->ec = NULL; } s->prevFrame = NULL; }   /* :81-84 */

/* ScreenPressor.hx:96-101 */
int sp_is_key(const uint8_t *d, int len)
{
    if (!d || len == 0) return 0;
    int b = d[0];
    return b == 0x12 || b == 0x11 || b == 0x22 || b == 0x21 || b == 0x32 || b == 0x31;
}

/* ScreenPressor.hx:66-79 */
static int init_entro(sp_dec *s, int version)
{
    switch (version) {
    case 2: s->ec = entro_rc_new(); break;
    case 3: s->ec = entro_ans_new(64); s->SC_CXSHIFT = 2; break;
    case 4: s->ec = entro_ans_new(32); s->SC_CXSHIFT = 2; break;
    default: return 0;
    }
    if (!s->ec) return 0;
    s->decodingBools = s->ec->canDecodeBool(s->ec);
    s->ec->preinit(s->ec);
    return 1;
}

/* ScreenPressor.hx:108-115 */
static int renew_i(sp_dec *s)
{
    s->prevFrame = NULL;
    if (s->last_one_was_flat >= 0) return 1;
    if (!s->ec) return 0;
    s->ec->renewI(s->ec);
    return 1;
}

#define RD(i) ((i) < len ? src[i] : 0)

static inline int32_t px_get(const int32_t *p, long i, long end) { return (i >= 0 && i < end) ? p[i] : 0; }

/* the three colour symbols of one pixel, ScreenPressor.hx:173-183 / :224-234 / :419-429 */
/* Defined behaviour (not in the reference): zero-length runs are legal syntax and cost almost no bits once their model
 * has adapted, so a hostile stream can keep a decoder busy for ever.  No encoder emits them; a frame that needs more
 * run-loop iterations than 2 per pixel + 16 per block + 4096 is reported as failed (ctx_fail doubles as the flag). */
#define SP_RUN_BUDGET(X, Y) (2L * (X) * (Y) + 16L * (((X) + 15) / 16) * (((Y) + 15) / 16) + 4096)
#define SP_SPEND(s) do { if (--(s)->budget < 0) { (s)->ctx_fail = 1; (s)->ec->fail((s)->ec); } } while (0)

/* Defined behaviour (not in the reference): cx + cx1 stays below 4096 on every valid stream (cx < 64, cx1 <= 0xFC0;
 * 16 bpp v2: 5-bit channel values).  A corrupt stream can exceed it -- the reference would read outside cntab[] --
 * so the index wraps inside its channel and the frame is reported as failed. */
/* test hook: when set, every colour-context index used is appended here (single-threaded callers only) */
int32_t *g_ora_ctx_trace = 0; long g_ora_ctx_trace_cap = 0, g_ora_ctx_trace_n = 0;
void ora_ctx_trace(int32_t *buf, long cap) { g_ora_ctx_trace = buf; g_ora_ctx_trace_cap = cap; g_ora_ctx_trace_n = 0; }
long ora_ctx_trace_count(void) { return g_ora_ctx_trace_n; }

static inline int ctx_index(sp_dec *s, int channel)
{
    int i = s->cx + s->cx1;
    if (i < 0 || i >= CC_CXMAX) { s->ctx_fail = 1; s->ec->fail(s->ec); i &= CC_CXMAX - 1; }
    if (g_ora_ctx_trace && g_ora_ctx_trace_n < g_ora_ctx_trace_cap) g_ora_ctx_trace[g_ora_ctx_trace_n++] = channel * CC_CXMAX + i;
    return channel * CC_CXMAX + i;
}

static inline int32_t decode_rgb(sp_dec *s)
{
    entro *ec = s->ec;
    int r = ec->decodeClr(ec, ctx_index(s, 0));
    s->cx1 = (s->cx << 6) & 0xFC0; s->cx = r >> s->SC_CXSHIFT;
    int g = ec->decodeClr(ec, ctx_index(s, 1));
    s->cx1 = (s->cx << 6) & 0xFC0; s->cx = g >> s->SC_CXSHIFT;
    int b = ec->decodeClr(ec, ctx_index(s, 2));
    s->cx1 = (s->cx << 6) & 0xFC0; s->cx = b >> s->SC_CXSHIFT;
    return (b << 16) + (g << 8) + r;
}

static inline int32_t grad(int32_t left, int32_t above, int32_t aboveleft)
{   /* per byte: left + above - aboveleft, & 0xFF (ScreenPressor.hx:261-264) */
    int r = (left & 0xFF) + (above & 0xFF) - (aboveleft & 0xFF);
    int g = ((left >> 8) & 0xFF) + ((above >> 8) & 0xFF) - ((aboveleft >> 8) & 0xFF);
    int b = ((left >> 16) & 0xFF) + ((above >> 16) & 0xFF) - ((aboveleft >> 16) & 0xFF);
    return ((b & 0xFF) << 16) + ((g & 0xFF) << 8) + (r & 0xFF);
}

/* ScreenPressor.hx:117-295 */
int sp_decompress_i(sp_dec *s, const uint8_t *src, int len, int32_t *dst)
{
    const int X = s->X;
    const long end = (long)X * s->Y;
    long di = 0, lasti = 0;
    int32_t clr = 0;
    int maskcx1 = 0xFC00, shiftcx1 = 4, shiftcx = 18;
    if (len <= 0) return ORA_ERROR_OCCURED;
    int head = src[0];
    int version = (head >> 4) + 1;
    if ((head & 0xF) == 1) {                                  /* flat, :132-155 */
        if (!s->ec && s->last_one_was_flat < 0) return ORA_ERROR_OCCURED;   /* ec == null dereference in the reference */
        renew_i(s);
        int32_t c;
        if (s->bpp == 16) {
            int clr16 = len >= 2 ? RD(0) + RD(1) * 256 : 0;               /* src[1] undefined -> NaN -> 0 under every `&` */
            int b = (clr16 & 0x1F) << 3, g = ((clr16 >> 5) & 0x1F) << 3, r = ((clr16 >> 10) & 0x1F) << 3;
            c = (r << 16) + (g << 8) + b;
        } else {
            c = len >= 2 ? (RD(3) << 16) + (RD(2) << 8) + RD(1) : 0;      /* `+ b` with b undefined is NaN -> stored as 0 */
        }
        for (long i = 0; i < end; i++) dst[i] = c;
        s->prevFrame = dst; s->last_one_was_flat = c; s->decodedI = 1;
        return ORA_ZERO_STATE;
    }
    s->last_one_was_flat = -1;
    if ((head & 0xF) != 2) return ORA_ERROR_OCCURED;          /* :157-159 */
    if (!s->ec && !init_entro(s, version)) return ORA_ERROR_OCCURED;
    renew_i(s);
    entro *ec = s->ec;
    ec->decodeBegin(ec, src, len, 1);
    /* the reference always runs DecompressI to its end (:293); a frame that FAILS here (defined behaviour) still
     * counts as "an I frame has been seen", so later P frames are decoded -- from nothing (prevFrame is null) */
    s->decodedI = 1;
    s->ctx_fail = 0;                                          /* the failure report is per frame */
    s->budget = SP_RUN_BUDGET(s->X, s->Y);
    s->cx = s->cx1 = 0;
    int k = 0;
    lasti = di;
    while (k < X + 1) {                                       /* :170-197 */
        SP_SPEND(s);
        clr = decode_rgb(s);
        int n = ec->decodeN(ec, 0);
        if ((ec->failed(ec) || s->ctx_fail)) return ORA_ERROR_OCCURED;
        k += n;
        while (n-- > 0) { if (di < end) dst[di] = clr; di++; }
        lasti = di - 1;
    }
    if (s->bpp == 16 && ec->differentConstantsFor16bpp(ec)) { maskcx1 = 0xFF00; shiftcx1 = 2; shiftcx = 16; }   /* :200-202 */
    const long off = -X - 1;
    int ptype = 0;
    while (di < end) {                                        /* :218-286 */
        SP_SPEND(s);
        ptype = ec->decodeP(ec, ptype);
        if (ptype == 0) clr = decode_rgb(s);
        int n = ec->decodeN(ec, ptype);
        if ((ec->failed(ec) || s->ctx_fail)) return ORA_ERROR_OCCURED;
        switch (ptype) {
        case 0:
            while (n-- > 0) { if (di < end) dst[di] = clr; di++; }
            lasti = di - 1; break;
        case 1:
            while (n-- > 0) { int32_t v = px_get(dst, lasti, end); if (di < end) dst[di] = v; lasti = di; di++; }
            clr = px_get(dst, lasti, end); break;
        case 2:
            while (n-- > 0) { clr = px_get(dst, di + off + 1, end); if (di < end) dst[di] = clr; di++; }
            lasti = di - 1; break;
        case 4:
            while (n-- > 0) {
                clr = grad(px_get(dst, lasti, end), px_get(dst, di + off + 1, end), px_get(dst, di + off, end));
                if (di < end) dst[di] = clr;
                lasti = di; di++;
            }
            break;
        case 5:
            while (n-- > 0) { clr = px_get(dst, di + off, end); if (di < end) dst[di] = clr; di++; }
            lasti = di - 1; break;
        default: break;                                       /* ptype 3 does nothing in an I frame */
        }
        s->cx1 = (clr & maskcx1) >> shiftcx1;                 /* :274-275 */
        s->cx = clr >> shiftcx;
    }
    s->prevFrame = dst;
    s->decodedI = 1;
    return ORA_ZERO_STATE;
}

/* ScreenPressor.hx:302-484 */
int sp_decompress_p(sp_dec *s, const uint8_t *src, int len, int32_t *dst, const int32_t **data_pnt, int *signif_out)
{
    s->last_one_was_flat = -1;
    *data_pnt = s->prevFrame; *signif_out = 0;
    if (len == 0 || !s->decodedI) return ORA_ZERO_STATE;      /* :308-309 */
    if (src[0] == 0) return ORA_ZERO_STATE;                   /* :311-313 */
    const int X = s->X, Y = s->Y, nbx = s->nbx, nby = s->nby;
    const long end = (long)X * Y;
    const int32_t *prev = s->prevFrame;
    int maskcx1 = 0xFC00, shiftcx1 = 4, shiftcx = 18;
    entro *ec = s->ec;
    if (ec->differentConstantsFor16bpp(ec) && s->bpp == 16) { maskcx1 = 0xFF00; shiftcx1 = 2; shiftcx = 16; }
    ec->decodeBegin(ec, src, len, 1);
    s->ctx_fail = 0;                                          /* the failure report is per frame */
    s->budget = SP_RUN_BUDGET(s->X, s->Y);
    int t = ec->decodeX(ec);
    int xx1 = ec->decodeX(ec); xx1 = (xx1 << 8) + t;
    t = ec->decodeX(ec);
    int xx2 = ec->decodeX(ec); xx2 = (xx2 << 8) + t;
    const int nb = nbx * nby;
    for (int i = 0; i < nb; i++) s->bts[i] = 0;
    long x = xx1;
    while (x <= xx2) {                                        /* :336-344 */
        SP_SPEND(s);
        int bt = ec->decodeBT(ec);
        int n = ec->decodeBN(ec);
        if ((ec->failed(ec) || s->ctx_fail)) return ORA_ERROR_OCCURED;
        for (int i = 0; i < n; i++) { if (x >= 0 && x < nb) s->bts[x] = bt; x++; }
    }
    int signif = 0;
    for (int i = s->insignificant_blocks > 0 ? s->insignificant_blocks : 0; i < nb; i++)
        if (s->bts[i] > 0) { signif = 1; break; }
    /* defined behaviour: start from the previous picture (see header) */
    if (prev) memcpy(dst, prev, (size_t)end * 4); else memset(dst, 0, (size_t)end * 4);
    const long off = -X - 1;
    int32_t clr = 0;
    s->cx = s->cx1 = 0;
    int lastmx = 0, lastmy = 0;
    for (int by = 0; by < nby; by++)
        for (int bx = 0; bx < nbx; bx++) {
            const int y16 = by * 16, x16 = bx * 16;
            int x1 = x16, x2 = x16 + 16, y1 = y16, y2 = y16 + 16;
            if (x2 > X) x2 = X;
            if (y2 > Y) y2 = Y;
            const int bt = s->bts[by * nbx + bx];
            if (bt <= 0) continue;                            /* copied from prev already */
            if (((bt - 1) & 1) > 0) {                         /* sub-rectangle, :375-386 */
                x1 = ec->decodeSXY(ec, 0) + x16;
                y1 = ec->decodeSXY(ec, 1) + y16;
                x2 = ec->decodeSXY(ec, 2) + x16 + 1;
                y2 = ec->decodeSXY(ec, 3) + y16 + 1;
            }
            if (((bt - 1) & 2) > 0) {                         /* motion vector, :388-405 */
                int mx, my;
                if (s->decodingBools && ec->decodeBool(ec)) { mx = lastmx; my = lastmy; }
                else { mx = ec->decodeMX(ec) - SP_MSR_X; my = ec->decodeMY(ec) - SP_MSR_Y; }
                if ((ec->failed(ec) || s->ctx_fail)) return ORA_ERROR_OCCURED;
                lastmx = mx; lastmy = my;
                for (int y = y1; y < y2; y++) {
                    long i = (long)y * X + x1, j = (long)(y + my) * X + (x1 + mx);
                    for (int xx = 0; xx < x2 - x1; xx++)
                        if (i + xx >= 0 && i + xx < end) dst[i + xx] = prev ? px_get(prev, j + xx, end) : 0;
                }
            } else {                                          /* data, :406-467 */
                int xq = x1, y = y1;
                int ptype = 0;
                while (y < y2) {
                    SP_SPEND(s);
                    long i = (long)y * X + xq;
                    ptype = ec->decodeP(ec, ptype);
                    if (ptype == 0) clr = decode_rgb(s);
                    int n = ec->decodeN(ec, ptype);
                    if ((ec->failed(ec) || s->ctx_fail)) return ORA_ERROR_OCCURED;
                    for (int c = 0; c < n; c++) {
                        switch (ptype) {
                        case 1: clr = px_get(dst, i - 1, end); break;
                        case 2: clr = px_get(dst, i + off + 1, end); break;
                        case 3: clr = prev ? px_get(prev, i, end) : 0; break;
                        /* JavaScript: dstbytes[k] with k < 0 is `undefined`, the sum is NaN and NaN & 0xFF is 0 -- ONE
                         * neighbour outside the buffer zeroes the whole pixel (ScreenPressor.hx:443-449); the above-left
                         * neighbour has the lowest index of the three.  Found by the second reading (oracle/sp_naive.py). */
                        case 4: clr = (i + off < 0) ? 0 : grad(px_get(dst, i - 1, end), px_get(dst, i + off + 1, end), px_get(dst, i + off, end)); break;
                        case 5: clr = px_get(dst, i + off, end); break;
                        default: break;
                        }
                        if (i >= 0 && i < end) dst[i] = clr;
                        xq++;
                        if (xq >= x2) { xq = x1; y++; i = (long)y * X + xq; } else i += 1;
                    }
                    s->cx1 = (clr & maskcx1) >> shiftcx1;     /* :462-463 */
                    s->cx = clr >> shiftcx;
                }
            }
        }
    /* a symbol that failed after the last in-loop check (e.g. the sub-rectangle of a block that then holds no rows) */
    if ((ec->failed(ec) || s->ctx_fail)) return ORA_ERROR_OCCURED;
    s->prevFrame = dst;
    *data_pnt = dst; *signif_out = signif;
    return ORA_ZERO_STATE;
}
