/*
 * msv1_synth.c -- synthetic, always-valid Microsoft Video 1 bitstreams.
 *
 * No real video is available offline, so tests and bench.py decode streams
 * produced here.  The opcode space follows the reference decoder's class tests
 * (reference src/MSVideo1.hx:131-181 for RGB555, :313-364 for 8-bit) and the
 * constraints a valid encoder must honour:
 *   16-bit 1-colour : word bit15 = 1 and high byte not in 0x84..0x87 (red != 1)
 *   16-bit 2/8-col. : stored flag word high byte < 0x80; colour0 bit15 = 0 / 1
 *   8-bit  2-colour : stored flag word high byte < 0x80 and word != 0 (terminator)
 *   8-bit  8-colour : stored flag word high byte >= 0x90
 *   8-bit  1-colour : high byte 0x80, low byte = palette index
 *   skip run        : 1 <= n <= 1023, may span block rows
 * This is a stream generator (random block classes / colours / flags drawn from
 * a recipe), not an image encoder: the decoder's work depends only on the
 * opcode mix, which the recipe controls exactly.
 */
#include <stdint.h>
#include <stddef.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r)
{   /* splitmix64 */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t rng_u32(rng_t *r) { return (uint32_t)(rng_next(r) >> 32); }
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((uint64_t)rng_u32(r) * n) >> 32); }

typedef struct {
    int32_t skip_start_permille; /* chance (per 1000) that a block starts a skip run; 0 => key frame */
    int32_t mean_skip;           /* mean run length (geometric), clipped to 1..1023 */
    int32_t pct1, pct2, pct8;    /* mix of coded blocks, percent, sums to 100 */
} jsp_msv1_recipe;

static inline uint16_t rgb555_for_1colour(rng_t *r)
{
    uint32_t red;
    do red = rng_below(r, 32); while (red == 1);       /* high byte 0x84..0x87 is the skip class */
    return (uint16_t)(0x8000u | (red << 10) | rng_below(r, 1024));
}

/* Returns the number of bytes written, or 0 if cap is too small. */
size_t jsp_synth_msv1_frame(int is8, int width, int height, uint64_t seed,
                            const jsp_msv1_recipe *rc, uint8_t *out, size_t cap)
{
    rng_t r = { seed * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull };
    const int nblocks = (width >> 2) * (height >> 2);
    size_t n = 0;
    int blk = 0;
    while (blk < nblocks) {
        if (n + 18 > cap) return 0;
        if (rc->skip_start_permille > 0 && (int)rng_below(&r, 1000) < rc->skip_start_permille) {
            /* geometric run length with the requested mean */
            int run = 1;
            uint32_t p = rc->mean_skip > 1 ? (uint32_t)(4294967296.0 / rc->mean_skip) : 0xFFFFFFFFu;
            while (run < 1023 && rng_u32(&r) >= p) run++;
            if (run > nblocks - blk) run = nblocks - blk;
            out[n++] = (uint8_t)(run & 0xFF);
            out[n++] = (uint8_t)(0x84 + (run >> 8));
            blk += run;
            continue;
        }
        int pick = (int)rng_below(&r, 100);
        int cls = pick < rc->pct1 ? 1 : (pick < rc->pct1 + rc->pct2 ? 2 : 8);
        if (is8) {
            if (cls == 1) {
                out[n++] = (uint8_t)rng_below(&r, 256);
                out[n++] = 0x80;
            } else if (cls == 2) {
                uint32_t flags;
                do flags = rng_below(&r, 0x8000); while (flags == 0);
                out[n++] = (uint8_t)flags; out[n++] = (uint8_t)(flags >> 8);
                out[n++] = (uint8_t)rng_below(&r, 256);
                out[n++] = (uint8_t)rng_below(&r, 256);
            } else {
                uint32_t flags;
                do flags = rng_below(&r, 0x10000); while ((flags >> 8) < 0x90);
                out[n++] = (uint8_t)flags; out[n++] = (uint8_t)(flags >> 8);
                for (int k = 0; k < 8; k++) out[n++] = (uint8_t)rng_below(&r, 256);
            }
        } else {
            if (cls == 1) {
                uint16_t c = rgb555_for_1colour(&r);
                out[n++] = (uint8_t)c; out[n++] = (uint8_t)(c >> 8);
            } else {
                uint32_t flags = rng_below(&r, 0x8000);
                out[n++] = (uint8_t)flags; out[n++] = (uint8_t)(flags >> 8);
                int ncol = cls == 2 ? 2 : 8;
                for (int k = 0; k < ncol; k++) {
                    uint32_t c = rng_below(&r, 0x10000);      /* bit15 of colours 1.. is ignored by the decoder */
                    if (k == 0) c = cls == 2 ? (c & 0x7FFF) : (c | 0x8000);
                    out[n++] = (uint8_t)c; out[n++] = (uint8_t)(c >> 8);
                }
            }
        }
        blk++;
    }
    return n;
}

/* Worst-case size of one frame produced above. */
size_t jsp_synth_msv1_bound(int is8, int width, int height)
{
    return (size_t)(width >> 2) * (height >> 2) * (is8 ? 10 : 18) + 32;
}
