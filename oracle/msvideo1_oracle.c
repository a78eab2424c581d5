/*
 * msvideo1_oracle.c -- CPU restatement of reference src/MSVideo1.hx
 * (MSVideo1_16bit :8-260, MSVideo1_8bit :262-429).  TEST INFRASTRUCTURE ONLY.
 *
 * JavaScript semantics made explicit:
 *  - an out-of-bounds Uint8Array read yields `undefined`; the arithmetic the
 *    reference does on it is spelled out at each site (UNDEF below);
 *  - `src[i] + src[i+1]*256` with an undefined operand is NaN, and every
 *    `NaN & mask` is 0, so a 16-bit word that straddles the end reads as 0;
 *  - storing undefined/NaN into an Int32Array stores 0.
 * Defined behaviour where the reference has none:
 *  - skip blocks before any frame exists copy zeros (the reference throws a
 *    TypeError on `prevFrame[..]` of null, MSVideo1.hx:79);
 *  - blocks left untouched by the 8-bit `00 00` terminator (MSVideo1.hx:313)
 *    take the previous picture (zeros if none); the reference leaves whatever
 *    the caller's buffer held.
 */
#include "oracle_internal.h"
#include <stdlib.h>
#include <string.h>

#define UNDEF (-1)

struct msv1_dec {
    int is8;
    int X, Y;
    int insignificant_blocks;   /* MSVideo1.hx:39,290 */
    int insign_lines;           /* MSVideo1.hx:40; never set by the 8-bit Preinit (:281-291) */
    int insign_lines_set;
    unsigned size_of_just_skips;/* MSVideo1.hx:29-30 */
    const int32_t *prevFrame;
    int32_t pal[256];
    uint8_t *block_changes;
    uint8_t *pal8; int pal8_len;
};

static inline int rd(const uint8_t *s, int len, int i) { return (i >= 0 && i < len) ? s[i] : UNDEF; }
/* src[si] + src[si+1]*256, NaN -> all masks give 0 (MSVideo1.hx:137,217) */
static inline int w16(const uint8_t *s, int len, int i) { return (i + 1 < len) ? (s[i] | (s[i + 1] << 8)) : 0; }
/* MSVideo1.hx:211-214 */
static inline int32_t fromRGB15(int c) { return ((c & 0x1F) << 3) + ((c & 0x3E0) << 6) + ((c & 0x7C00) << 9); }

msv1_dec *msv1_new(int is8, int w, int h, const uint8_t *palette, int palette_bytes)
{
    msv1_dec *m = (msv1_dec *)calloc(1, sizeof *m);
    m->is8 = is8; m->X = w; m->Y = h;
    int nby = h >> 2;
    m->block_changes = (uint8_t *)calloc(nby > 0 ? nby : 1, 1);
    unsigned nblocks = (unsigned)((w >> 2) * (h >> 2));
    m->size_of_just_skips = nblocks / 1023 * 2 + 10;       /* :29-30 */
    if (is8 && palette && palette_bytes > 0) {
        m->pal8 = (uint8_t *)malloc(palette_bytes);
        memcpy(m->pal8, palette, palette_bytes);
        m->pal8_len = palette_bytes;
    }
    return m;
}

void msv1_free(msv1_dec *m) { if (!m) return; free(m->block_changes); free(m->pal8); free(m); }

void msv1_preinit(msv1_dec *m, int insignificant_lines)
{
    if (m->is8) {
        /* MSVideo1.hx:281-291: up to 256 little-endian u32 (B,G,R,reserved); insign_lines NOT set */
        int i = 0, pos = 0;
        while (i < 256 && m->pal8_len - pos >= 4) {
            const uint8_t *p = m->pal8 + pos;
            m->pal[i] = (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
            pos += 4; i++;
        }
        m->insignificant_blocks = (insignificant_lines + 3) >> 2;
    } else {
        m->insignificant_blocks = (insignificant_lines + 3) >> 2;  /* :37-41 */
        m->insign_lines = insignificant_lines;
        m->insign_lines_set = 1;
    }
}

const int32_t *msv1_prev(msv1_dec *m) { return m->prevFrame; }

/* MSVideo1.hx:74-84 */
static inline void copy_block(msv1_dec *m, int di, int32_t *dst)
{
    const int32_t *pv = m->prevFrame;
    for (int y = 0; y < 4; y++) {
        if (pv) memcpy(dst + di, pv + di, 16); else memset(dst + di, 0, 16);
        di += m->X;
    }
}

static inline void fill1(int32_t *dst, int di, int X, int32_t c)
{
    for (int y = 0; y < 4; y++) { dst[di] = c; dst[di + 1] = c; dst[di + 2] = c; dst[di + 3] = c; di += X; }
}

/* MSVideo1.hx:86-104 */
static int just_skip_blocks(msv1_dec *m, const uint8_t *src, int len)
{
    int si = 0, n = 0, nblocks = (m->X >> 2) * (m->Y >> 2);
    while (si < len) {
        int a = src[si];
        int b = rd(src, len, si + 1);
        if (b != UNDEF && (b & 0xFC) == 0x84) {
            n += ((b - 0x84) << 8) + a;
            if (n >= nblocks) return 1;
        } else
            return 0;
        si += 2;
    }
    return 1;
}

static int significance(msv1_dec *m, int changes, const int32_t *dst)
{
    int nby = m->Y >> 2;
    int signif = 0;
    if (changes)
        for (int i = m->insignificant_blocks > 0 ? m->insignificant_blocks : 0; i < nby; i++)
            if (m->block_changes[i]) { signif = 1; break; }
    if (signif && m->prevFrame) {
        signif = 0;
        /* 8-bit: insign_lines is undefined in JS => `NaN < Y*X` is false and the compare
         * loop never runs (MSVideo1.hx:380-388 vs :281-291) */
        if (m->insign_lines_set) {
            long n = (long)m->Y * m->X;
            for (long i = (long)m->insign_lines * m->X; i < n; i++)
                if (i >= 0 && dst[i] != m->prevFrame[i]) { signif = 1; break; }
        }
    }
    return signif;
}

/* MSVideo1.hx:106-209 */
static void decompress_p16(msv1_dec *m, const uint8_t *src, int len, int32_t *dst,
                           const int32_t **data_pnt, int *signif_out)
{
    if (len == 0 || ((unsigned)len < m->size_of_just_skips && just_skip_blocks(m, src, len))) {
        *data_pnt = m->prevFrame; *signif_out = 0; return;
    }
    const int X = m->X, nbx = X >> 2, nby = m->Y >> 2;
    int skip = 0, si = 0, changes = 0;
    int32_t pal[8];
    for (int by = 0; by < nby; by++) {
        int di = by * X * 4;
        m->block_changes[by] = 0;
        for (int bx = 0; bx < nbx; bx++, di += 4) {
            if (skip != 0) { skip--; copy_block(m, di, dst); continue; }
            int a = rd(src, len, si), b = rd(src, len, si + 1);
            si += 2;
            if (b != UNDEF && (b & 0xFC) == 0x84) {
                skip = ((b - 0x84) << 8) + a - 1;
                copy_block(m, di, dst);
            } else if (b != UNDEF && b < 0x80) {
                int flags = ((b << 8) + a) ^ 0xFFFF;
                int clr0 = w16(src, len, si);
                pal[0] = fromRGB15(clr0);
                pal[1] = fromRGB15(w16(src, len, si + 2));
                si += 4;
                if (clr0 & 0x8000) {
                    for (int k = 0; k < 6; k++) pal[2 + k] = fromRGB15(w16(src, len, si + 2 * k));
                    si += 12;
                    int d = di;
                    for (int y = 0; y < 4; y++) {
                        int ty = (y & 2) << 1;
                        for (int x = 0; x < 4; x++) { dst[d + x] = pal[ty + (x & 2) + (flags & 1)]; flags >>= 1; }
                        d += X;
                    }
                } else {
                    int d = di;
                    for (int y = 0; y < 4; y++) {
                        for (int x = 0; x < 4; x++) { dst[d + x] = pal[flags & 1]; flags >>= 1; }
                        d += X;
                    }
                }
                changes = 1; m->block_changes[by] = 1;
            } else {
                /* b >= 0x80, or b undefined: (undefined<<8)+a = a; a undefined too => NaN => 0 */
                int w = (b == UNDEF) ? (a == UNDEF ? 0 : a) : ((b << 8) + a);
                fill1(dst, di, X, fromRGB15(w));
                changes = 1; m->block_changes[by] = 1;
            }
        }
    }
    int signif = significance(m, changes, dst);
    if (changes) m->prevFrame = dst;
    *data_pnt = m->prevFrame; *signif_out = signif;
}

/* MSVideo1.hx:293-393 */
static void decompress_p8(msv1_dec *m, const uint8_t *src, int len, int32_t *dst,
                          const int32_t **data_pnt, int *signif_out)
{
    const int X = m->X, nbx = X >> 2, nby = m->Y >> 2;
    int skip = 0, si = 0, changes = 0, stopped = 0;
    int32_t p2[8];
    const int32_t *pal = m->pal;
#define PAL(ix) ((ix) == UNDEF ? 0 : pal[ix])
    for (int by = 0; by < nby; by++) {
        int di = by * X * 4;
        if (!stopped) m->block_changes[by] = 0;   /* rows after the terminator keep their old flag (:305 not reached) */
        for (int bx = 0; bx < nbx; bx++, di += 4) {
            if (stopped) { copy_block(m, di, dst); continue; }   /* defined behaviour, see header */
            if (skip != 0) { skip--; copy_block(m, di, dst); continue; }
            int a = rd(src, len, si), b = rd(src, len, si + 1);
            if (a != UNDEF && b != UNDEF && a + b == 0) {         /* :313 `throw 0` */
                stopped = 1; copy_block(m, di, dst); continue;
            }
            si += 2;
            if (b != UNDEF && (b & 0xFC) == 0x84) {
                skip = ((b - 0x84) << 8) + a - 1;
                copy_block(m, di, dst);
            } else if (b != UNDEF && b < 0x80) {
                int flags = (b << 8) + a;
                p2[1] = PAL(rd(src, len, si));
                p2[0] = PAL(rd(src, len, si + 1));
                si += 2;
                int d = di;
                for (int y = 0; y < 4; y++) {
                    for (int x = 0; x < 4; x++) { dst[d + x] = p2[flags & 1]; flags >>= 1; }
                    d += X;
                }
                changes = 1; m->block_changes[by] = 1;
            } else if (b != UNDEF && b >= 0x90) {
                int flags = ((b << 8) + a) ^ 0xFFFF;
                for (int i = 0; i < 8; i++) p2[i] = PAL(rd(src, len, si + i));
                si += 8;
                int d = di;
                for (int y = 0; y < 4; y++) {
                    int ty = (y & 2) << 1;
                    for (int x = 0; x < 4; x++) { dst[d + x] = p2[ty + (x & 2) + (flags & 1)]; flags >>= 1; }
                    d += X;
                }
                changes = 1; m->block_changes[by] = 1;
            } else {
                fill1(dst, di, X, PAL(a));
                changes = 1; m->block_changes[by] = 1;
            }
        }
    }
#undef PAL
    int signif = significance(m, changes, dst);
    if (changes) m->prevFrame = dst;
    *data_pnt = m->prevFrame; *signif_out = signif;
}

void msv1_decompress_p(msv1_dec *m, const uint8_t *src, int len, int32_t *dst,
                       const int32_t **data_pnt, int *signif)
{
    if (m->is8) decompress_p8(m, src, len, dst, data_pnt, signif);
    else decompress_p16(m, src, len, dst, data_pnt, signif);
}

/* MSVideo1.hx:226-259 (16-bit) and :395-427 (8-bit) */
int msv1_is_key(msv1_dec *m, const uint8_t *src, int len)
{
    if (len == 0) return 0;
    const int nbx = m->X >> 2, nby = m->Y >> 2;
    int skip = 0, si = 0, key = 1;
    for (int by = 0; by < nby; by++)
        for (int bx = 0; bx < nbx; bx++) {
            if (skip != 0) { skip--; continue; }
            int a = rd(src, len, si), b = rd(src, len, si + 1);
            if (m->is8 && a != UNDEF && b != UNDEF && a + b == 0) return key;   /* :410 */
            si += 2;
            if (b != UNDEF && (b & 0xFC) == 0x84) {
                if (!m->is8) return 0;                                           /* :246 */
                skip = ((b - 0x84) << 8) + a - 1; key = 0;
            } else if (b != UNDEF && b < 0x80) {
                if (m->is8) si += 2;
                else si += (w16(src, len, si) & 0x8000) ? 16 : 4;
            } else if (m->is8 && b != UNDEF && b >= 0x90)
                si += 8;
        }
    return key;
}
