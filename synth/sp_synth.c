/*
 * sp_synth.c -- synthetic ScreenPressor encoder (I frames, flat frames, P frames) and a "screen content"
 * picture generator.  Produces exactly the syntax the reference decoder reads:
 *   I frame  src/ScreenPressor.hx:117-295   [head][payload]; first X+1 pixels as (colour, run), then
 *            (ptype in {0,1,2,4,5}, [colour], run) with runs <= 255
 *   P frame  src/ScreenPressor.hx:302-484   [changes][payload]; xx1/xx2, block-type RLE, then per changed
 *            16x16 block [sub-rectangle] [motion vector | runs with ptype 0..5]
 * Lossless, greedy longest-run choice among the predictors.  Predictions that would read outside the picture
 * are never chosen, so the streams do not depend on out-of-bounds behaviour.  Pixels are 0x00c3c2c1 with c1
 * the first coded channel (ScreenPressor.hx:189).
 */
#include "sp_coder.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
    int X, Y, bpp, version;
    int cxshift;                    /* SC_CXSHIFT, ScreenPressor.hx:59,71,73 */
    int maskcx1, shiftcx1, shiftcx; /* :122,200-202 */
    int cx, cx1;
    int have_i;
    sp_coder *ec;
    uint8_t *bts;
} sp_enc;

/* ext: an entropy coder supplied by the caller (the test suite passes its rANS coder for v3/v4), or NULL */
sp_enc *jsp_sp_enc_new_ext(int w, int h, int bpp, int version, sp_coder *ext)
{
    sp_enc *e = (sp_enc *)calloc(1, sizeof *e);
    e->X = w; e->Y = h; e->bpp = bpp; e->version = version;
    e->cxshift = (bpp == 16 && version == 2) ? 0 : 2;
    e->maskcx1 = 0xFC00; e->shiftcx1 = 4; e->shiftcx = 18;
    if (bpp == 16 && version == 2) { e->maskcx1 = 0xFF00; e->shiftcx1 = 2; e->shiftcx = 16; }
    e->ec = ext ? ext : (version == 2 ? sp_rc_coder_new() : sp_ans_coder_new(version == 3 ? 64 : 32));
    e->bts = (uint8_t *)calloc((size_t)((w + 15) / 16) * ((h + 15) / 16) + 1, 1);
    if (!e->ec) { free(e->bts); free(e); return NULL; }
    return e;
}

sp_enc *jsp_sp_enc_new(int w, int h, int bpp, int version) { return jsp_sp_enc_new_ext(w, h, bpp, version, NULL); }

void jsp_sp_enc_free(sp_enc *e) { if (!e) return; e->ec->destroy(e->ec); free(e->bts); free(e); }

static void put_rgb(sp_enc *e, int32_t clr)
{
    const int r = clr & 0xFF, g = (clr >> 8) & 0xFF, b = (clr >> 16) & 0xFF;
    e->ec->clr(e->ec, e->cx + e->cx1, r);
    e->cx1 = (e->cx << 6) & 0xFC0; e->cx = r >> e->cxshift;
    e->ec->clr(e->ec, 4096 + e->cx + e->cx1, g);
    e->cx1 = (e->cx << 6) & 0xFC0; e->cx = g >> e->cxshift;
    e->ec->clr(e->ec, 2 * 4096 + e->cx + e->cx1, b);
    e->cx1 = (e->cx << 6) & 0xFC0; e->cx = b >> e->cxshift;
}
static void ctx_from_pixel(sp_enc *e, int32_t clr)
{
    e->cx1 = (clr & e->maskcx1) >> e->shiftcx1;
    e->cx = clr >> e->shiftcx;
}
static inline int32_t grad(int32_t l, int32_t a, int32_t al)
{
    int r = (l & 0xFF) + (a & 0xFF) - (al & 0xFF);
    int g = ((l >> 8) & 0xFF) + ((a >> 8) & 0xFF) - ((al >> 8) & 0xFF);
    int b = ((l >> 16) & 0xFF) + ((a >> 16) & 0xFF) - ((al >> 16) & 0xFF);
    return ((b & 0xFF) << 16) | ((g & 0xFF) << 8) | (r & 0xFF);
}

size_t jsp_sp_enc_flat(sp_enc *e, int32_t colour, uint8_t *out, size_t cap)
{   /* ScreenPressor.hx:132-155 (24/32 bpp form) */
    if (cap < 4) return 0;
    out[0] = (uint8_t)(((e->version - 1) << 4) | 1);
    out[1] = colour & 0xFF; out[2] = (colour >> 8) & 0xFF; out[3] = (colour >> 16) & 0xFF;
    if (e->have_i) e->ec->renew_i(e->ec);      /* the decoder resets its models here too (RenewI, :108-115) */
    return 4;
}

size_t jsp_sp_enc_iframe(sp_enc *e, const int32_t *px, uint8_t *out, size_t cap)
{
    const int X = e->X; const long end = (long)X * e->Y;
    sp_coder *ec = e->ec;
    if (cap < 16) return 0;
    out[0] = (uint8_t)(((e->version - 1) << 4) | 2);
    ec->renew_i(ec); ec->begin(ec);
    e->have_i = 1;
    e->cx = e->cx1 = 0;
    long di = 0;
    int32_t clr = 0;
    while (di < X + 1 && di < end) {                      /* first X+1 pixels: (colour, run) */
        clr = px[di];
        int n = 1;
        while (n < 255 && di + n < end && px[di + n] == clr) n++;
        put_rgb(e, clr);
        ec->n(ec, 0, n);
        di += n;
    }
    int ptype = 0;
    while (di < end) {
        const long rem = end - di;
        const int lim = rem < 255 ? (int)rem : 255;
        int best = 0, bestn = 0, n;
        /* 1: repeat the previous pixel */
        for (n = 0; n < lim && px[di + n] == px[di - 1]; n++) {}
        if (n > bestn) { bestn = n; best = 1; }
        /* 2: copy from above */
        for (n = 0; n < lim && px[di + n] == px[di + n - X]; n++) {}
        if (n > bestn) { bestn = n; best = 2; }
        /* 5: copy from above-left */
        for (n = 0; n < lim && px[di + n] == px[di + n - X - 1]; n++) {}
        if (n > bestn) { bestn = n; best = 5; }
        /* 4: left + above - aboveleft */
        for (n = 0; n < lim && px[di + n] == grad(px[di + n - 1], px[di + n - X], px[di + n - X - 1]); n++) {}
        if (n > bestn) { bestn = n; best = 4; }
        if (bestn == 0) {
            best = 0; clr = px[di];
            for (bestn = 1; bestn < lim && px[di + bestn] == clr; bestn++) {}
        }
        ec->p(ec, ptype, best);
        ptype = best;
        if (best == 0) put_rgb(e, clr);
        ec->n(ec, best, bestn);
        di += bestn;
        ctx_from_pixel(e, px[di - 1]);
    }
    size_t n = ec->finish(ec, out + 1, cap - 1);
    return n ? n + 1 : 0;
}

/* mvx/mvy: a motion vector to try for changed blocks (0,0 = none). */
size_t jsp_sp_enc_pframe(sp_enc *e, const int32_t *px, const int32_t *prev, int mvx, int mvy, uint8_t *out, size_t cap)
{
    const int X = e->X, Y = e->Y, nbx = (X + 15) / 16, nby = (Y + 15) / 16, nb = nbx * nby;
    sp_coder *ec = e->ec;
    if (cap < 16 || !e->have_i) return 0;
    /* classify blocks */
    int first = -1, last = -1;
    int16_t *rect = (int16_t *)malloc((size_t)nb * 4 * sizeof(int16_t));
    for (int by = 0; by < nby; by++)
        for (int bx = 0; bx < nbx; bx++) {
            const int bi = by * nbx + bx, x16 = bx * 16, y16 = by * 16;
            const int xe = x16 + 16 > X ? X : x16 + 16, ye = y16 + 16 > Y ? Y : y16 + 16;
            int x1 = 99999, y1 = 99999, x2 = -1, y2 = -1;
            for (int y = y16; y < ye; y++)
                for (int x = x16; x < xe; x++)
                    if (px[(long)y * X + x] != prev[(long)y * X + x]) {
                        if (x < x1) x1 = x; if (x > x2) x2 = x; if (y < y1) y1 = y; if (y > y2) y2 = y;
                    }
            if (x2 < 0) { e->bts[bi] = 0; continue; }
            if (first < 0) first = bi;
            last = bi;
            const int full = (x1 == x16 && y1 == y16 && x2 == xe - 1 && y2 == ye - 1);
            /* does the motion vector reproduce the changed rectangle? (source must lie inside the picture) */
            int motion = 0;
            if (mvx || mvy) {
                const int rx1 = full ? x16 : x1, ry1 = full ? y16 : y1, rx2 = full ? xe : x2 + 1, ry2 = full ? ye : y2 + 1;
                if (rx1 + mvx >= 0 && rx2 + mvx <= X && ry1 + mvy >= 0 && ry2 + mvy <= Y) {
                    motion = 1;
                    for (int y = ry1; y < ry2 && motion; y++)
                        for (int x = rx1; x < rx2; x++)
                            if (px[(long)y * X + x] != prev[(long)(y + mvy) * X + x + mvx]) { motion = 0; break; }
                }
            }
            e->bts[bi] = (uint8_t)(1 + (full ? 0 : 1) + (motion ? 2 : 0));
            rect[bi * 4 + 0] = (int16_t)(full ? x16 : x1); rect[bi * 4 + 1] = (int16_t)(full ? y16 : y1);
            rect[bi * 4 + 2] = (int16_t)(full ? xe : x2 + 1); rect[bi * 4 + 3] = (int16_t)(full ? ye : y2 + 1);
        }
    if (first < 0) { free(rect); out[0] = 0; return 1; }          /* unchanged frame, ScreenPressor.hx:311-313 */
    out[0] = 1;
    ec->begin(ec);
    ec->x(ec, first & 0xFF); ec->x(ec, first >> 8); ec->x(ec, last & 0xFF); ec->x(ec, last >> 8);
    for (int x = first; x <= last;) {
        int n = 1;
        while (n < 255 && x + n <= last && e->bts[x + n] == e->bts[x]) n++;
        ec->bt(ec, e->bts[x]); ec->bn(ec, n);
        x += n;
    }
    e->cx = e->cx1 = 0;
    int lastmx = 0, lastmy = 0, have_last = 0;
    /* What the DECODER has in its destination while it decodes a block: the previous picture (ScreenPressor.hx:468-474 copies
     * unchanged blocks; DESIGN.md section 2 defines the start of a P frame as the previous picture) with the blocks decoded
     * so far replaced.  The predictors read neighbours from there -- at a picture's first column "left" and "above-left"
     * wrap to the previous row's LAST pixel, which belongs to a block decoded later and so still shows the previous frame. */
    int32_t *dv = (int32_t *)malloc((size_t)X * Y * sizeof(int32_t));
    memcpy(dv, prev, (size_t)X * Y * sizeof(int32_t));
    for (int bi = first; bi <= last; bi++) {
        const int bt = e->bts[bi];
        if (!bt) continue;
        const int bx = bi % nbx, by = bi / nbx, x16 = bx * 16, y16 = by * 16;
        const int x1 = rect[bi * 4], y1 = rect[bi * 4 + 1], x2 = rect[bi * 4 + 2], y2 = rect[bi * 4 + 3];
        for (int y = y1; y < y2; y++) memcpy(dv + (long)y * X + x1, px + (long)y * X + x1, (size_t)(x2 - x1) * sizeof(int32_t));
        if ((bt - 1) & 1) {
            ec->sxy(ec, 0, x1 - x16); ec->sxy(ec, 1, y1 - y16); ec->sxy(ec, 2, x2 - 1 - x16); ec->sxy(ec, 3, y2 - 1 - y16);
        }
        if ((bt - 1) & 2) {
            if (ec->can_bool(ec)) {
                const int same = have_last && lastmx == mvx && lastmy == mvy;
                ec->boolean(ec, same);
                if (!same) { ec->mx(ec, mvx + 256); ec->my(ec, mvy + 256); }
            } else { ec->mx(ec, mvx + 256); ec->my(ec, mvy + 256); }
            lastmx = mvx; lastmy = mvy; have_last = 1;
            continue;
        }
        /* data block: runs in raster order inside the rectangle */
        const int w = x2 - x1, total = w * (y2 - y1);
        int pos = 0, ptype = 0;
        int32_t clr = 0;
#define PIX(k) ((long)(y1 + (k) / w) * X + x1 + (k) % w)
        while (pos < total) {
            const int lim = total - pos < 255 ? total - pos : 255;
            int best = 0, bestn = 0, n;
            for (n = 0; n < lim; n++) { long i = PIX(pos + n); if (i < 1 || px[i] != dv[i - 1]) break; }
            if (n > bestn) { bestn = n; best = 1; }
            for (n = 0; n < lim; n++) { long i = PIX(pos + n); if (i - X < 0 || px[i] != dv[i - X]) break; }
            if (n > bestn) { bestn = n; best = 2; }
            for (n = 0; n < lim; n++) { long i = PIX(pos + n); if (px[i] != prev[i]) break; }
            if (n > bestn) { bestn = n; best = 3; }
            for (n = 0; n < lim; n++) { long i = PIX(pos + n); if (i - X - 1 < 0 || px[i] != grad(dv[i - 1], dv[i - X], dv[i - X - 1])) break; }
            if (n > bestn) { bestn = n; best = 4; }
            for (n = 0; n < lim; n++) { long i = PIX(pos + n); if (i - X - 1 < 0 || px[i] != dv[i - X - 1]) break; }
            if (n > bestn) { bestn = n; best = 5; }
            if (bestn == 0) {
                best = 0; clr = px[PIX(pos)];
                for (bestn = 1; bestn < lim && px[PIX(pos + bestn)] == clr; bestn++) {}
            }
            ec->p(ec, ptype, best);
            ptype = best;
            if (best == 0) put_rgb(e, clr);
            ec->n(ec, best, bestn);
            pos += bestn;
            ctx_from_pixel(e, px[PIX(pos - 1)]);
        }
#undef PIX
    }
    free(rect); free(dv);
    size_t n = ec->finish(ec, out + 1, cap - 1);
    return n ? n + 1 : 0;
}

/* ------------------------------------------------------------------------------------------------
 * "screen content" pictures (SURVEY.md 8d, C3/C4 recipe): flat background, solid rectangles, two-colour
 * "text" regions with 30 % ink, linear gradients (exercise predictor 4), a window that scrolls between
 * frames (motion blocks), small edits elsewhere (data blocks, sub-rectangles, predictor 3).
 * ------------------------------------------------------------------------------------------------ */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * n) >> 32); }
static inline int32_t rng_colour(rng_t *r, int bits) { int m = (1 << bits) - 1; return (int32_t)((rng_below(r, m + 1) << 16) | (rng_below(r, m + 1) << 8) | rng_below(r, m + 1)); }

static void fill_rect(int32_t *px, int X, int Y, int x0, int y0, int w, int h, int32_t c)
{
    for (int y = y0; y < y0 + h && y < Y; y++)
        for (int x = x0; x < x0 + w && x < X; x++) px[(long)y * X + x] = c;
}
static void text_rect(int32_t *px, int X, int Y, int x0, int y0, int w, int h, int32_t bg, int32_t ink, rng_t *r)
{
    for (int y = y0; y < y0 + h && y < Y; y++)
        for (int x = x0; x < x0 + w && x < X; x++) {
            const int in_line = ((y - y0) % 12) < 8;           /* text lines 8 px high, 4 px leading */
            px[(long)y * X + x] = (in_line && rng_below(r, 100) < 30) ? ink : bg;
        }
}
static void gradient_rect(int32_t *px, int X, int Y, int x0, int y0, int w, int h, int32_t c0, int dx, int dy, int maxv)
{
    for (int y = y0; y < y0 + h && y < Y; y++)
        for (int x = x0; x < x0 + w && x < X; x++) {
            int v = (x - x0) * dx + (y - y0) * dy;
            int r = ((c0 & 0xFF) + v) & maxv, g = (((c0 >> 8) & 0xFF) + v) & maxv, b = (((c0 >> 16) & 0xFF) + 2 * v) & maxv;
            px[(long)y * X + x] = (b << 16) | (g << 8) | r;
        }
}

/* bits = 8 for 24 bpp content, 5 for 16 bpp v2 content (channel values < 32) */
void jsp_synth_screen(int X, int Y, uint64_t seed, int bits, int32_t *px)
{
    rng_t r = { seed * 0xA24BAED4963EE407ull + 0x9FB21C651E98DF25ull };
    const int maxv = (1 << bits) - 1;
    int32_t bg = rng_colour(&r, bits);
    for (long i = 0; i < (long)X * Y; i++) px[i] = bg;
    int nrect = 20 + (int)rng_below(&r, 41);
    for (int k = 0; k < nrect; k++)
        fill_rect(px, X, Y, (int)rng_below(&r, X), (int)rng_below(&r, Y), 8 + (int)rng_below(&r, X / 3 + 1), 8 + (int)rng_below(&r, Y / 3 + 1), rng_colour(&r, bits));
    int ntext = 5 + (int)rng_below(&r, 11);
    for (int k = 0; k < ntext; k++)
        text_rect(px, X, Y, (int)rng_below(&r, X), (int)rng_below(&r, Y), 16 + (int)rng_below(&r, X / 3 + 1), 12 + (int)rng_below(&r, Y / 4 + 1), rng_colour(&r, bits), rng_colour(&r, bits), &r);
    int ngrad = 1 + (int)rng_below(&r, 3);
    for (int k = 0; k < ngrad; k++)
        gradient_rect(px, X, Y, (int)rng_below(&r, X), (int)rng_below(&r, Y), 16 + (int)rng_below(&r, X / 4 + 1), 16 + (int)rng_below(&r, Y / 4 + 1), rng_colour(&r, bits), 1 + (int)rng_below(&r, 2), (int)rng_below(&r, 2), maxv);
}

/* Next picture of a stream: a window scrolls by (mvx,mvy) (block aligned, so whole blocks match the motion
 * vector), a few small edits happen elsewhere.  change_permille ~ share of 16x16 blocks touched by edits. */
void jsp_synth_screen_next(int X, int Y, uint64_t seed, int bits, const int32_t *prev, int32_t *px,
                           int change_permille, int *mvx, int *mvy)
{
    rng_t r = { seed * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull };
    memcpy(px, prev, (size_t)X * Y * 4);
    *mvx = 0; *mvy = 0;
    const int nbx = X / 16, nby = Y / 16;
    if (nbx >= 6 && nby >= 6 && rng_below(&r, 100) < 70) {
        /* scrolling window: block-aligned region, content moves up by 1..8 pixels (my > 0 reads from below) */
        const int wx = 16 * (1 + (int)rng_below(&r, nbx / 2)), wy = 16 * (1 + (int)rng_below(&r, nby / 2));
        const int wbw = 1 + (int)rng_below(&r, nbx / 8 + 1), wbh = 1 + (int)rng_below(&r, nby / 8 + 1);
        const int my = 1 + (int)rng_below(&r, 8), mx = (int)rng_below(&r, 3) - 1;
        const int x0 = wx, y0 = wy, x1 = wx + 16 * wbw, y1 = wy + 16 * wbh;
        if (x1 + 2 <= X && y1 + my <= Y && x0 >= 2) {
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) px[(long)y * X + x] = prev[(long)(y + my) * X + x + mx];
            *mvx = mx; *mvy = my;
        }
    }
    const int nb = ((X + 15) / 16) * ((Y + 15) / 16);
    int nedit = (int)((long)nb * change_permille / 1000);
    if (nedit < 1) nedit = 1;
    for (int k = 0; k < nedit; k++) {
        const int x = (int)rng_below(&r, X), y = (int)rng_below(&r, Y);
        const int kind = (int)rng_below(&r, 3);
        if (kind == 0) fill_rect(px, X, Y, x, y, 1 + (int)rng_below(&r, 24), 1 + (int)rng_below(&r, 24), rng_colour(&r, bits));
        else if (kind == 1) text_rect(px, X, Y, x, y, 4 + (int)rng_below(&r, 28), 4 + (int)rng_below(&r, 12), rng_colour(&r, bits), rng_colour(&r, bits), &r);
        else px[(long)y * X + x] = rng_colour(&r, bits);
    }
}

/* Pictures that push the rANS colour contexts through all their kinds: vertical bands, band k drawn with random
 * pixels from a palette of ncolors[k] random colours (few colours -> small contexts, many -> 256-symbol tables). */
void jsp_synth_noise(int X, int Y, uint64_t seed, int bits, const int *ncolors, int nbands, int32_t *px)
{
    rng_t r = { seed * 0x9E3779B97F4A7C15ull + 0x1234567 };
    int32_t pal[256];
    for (int k = 0; k < nbands; k++) {
        int nc = ncolors[k] < 1 ? 1 : (ncolors[k] > 256 ? 256 : ncolors[k]);
        for (int i = 0; i < nc; i++) pal[i] = rng_colour(&r, bits);
        const int x0 = (int)((long)X * k / nbands), x1 = (int)((long)X * (k + 1) / nbands);
        for (int y = 0; y < Y; y++)
            for (int x = x0; x < x1; x++) px[(long)y * X + x] = pal[rng_below(&r, (uint32_t)nc)];
    }
}
