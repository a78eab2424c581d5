/*
 * oracle.c -- IVideoCodec-shaped dispatcher over the codec restatements, the
 * whole-stream helper and the multi-threaded CPU-baseline driver.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The frame loop follows Manager.worker (reference src/Manager.hx:454-525):
 * key frames go to DecompressI, the rest to DecompressP; the output buffer is
 * never the codec's PreviousFrame() (Manager.hx:470-477).
 */
#include "oracle_internal.h"
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

ora_dec *ora_create(int codec, int width, int height, int bpp, const uint8_t *palette, int palette_bytes)
{
    ora_dec *d = (ora_dec *)calloc(1, sizeof *d);
    d->codec = codec; d->X = width; d->Y = height; d->bpp = bpp;
    if (codec == ORA_CODEC_MSVC16) d->msv1 = msv1_new(0, width, height, NULL, 0);
    else if (codec == ORA_CODEC_MSVC8) d->msv1 = msv1_new(1, width, height, palette, palette_bytes);
    else d->sp = sp_new(width, height, bpp);
    return d;
}

void ora_destroy(ora_dec *d)
{
    if (!d) return;
    if (d->msv1) msv1_free(d->msv1);
    if (d->sp) sp_free(d->sp);
    free(d);
}

void ora_preinit(ora_dec *d, int n) { if (d->msv1) msv1_preinit(d->msv1, n); else sp_preinit(d->sp, n); }
int ora_is_key_frame(ora_dec *d, const uint8_t *data, int len)
{ return d->msv1 ? msv1_is_key(d->msv1, data, len) : sp_is_key(data, len); }
int ora_needs_index(ora_dec *d) { return d->msv1 ? 1 : 0; }  /* MSVideo1.hx:221-224, ScreenPressor.hx:486-489 */
const int32_t *ora_previous_frame(ora_dec *d) { return d->msv1 ? msv1_prev(d->msv1) : sp_prev(d->sp); }
int ora_state(ora_dec *d) { (void)d; return ORA_ZERO_STATE; }

int ora_decompress_i(ora_dec *d, const uint8_t *src, int len, int32_t *dst)
{
    if (d->msv1) {               /* MSVideo1.hx:62-67: DecompressI = DecompressP, result dropped */
        const int32_t *p; int s;
        msv1_decompress_p(d->msv1, src, len, dst, &p, &s);
        return ORA_ZERO_STATE;
    }
    return sp_decompress_i(d->sp, src, len, dst);
}

int ora_decompress_p(ora_dec *d, const uint8_t *src, int len, int32_t *dst,
                     const int32_t **data_pnt, int *significant)
{
    if (d->msv1) { msv1_decompress_p(d->msv1, src, len, dst, data_pnt, significant); return ORA_ZERO_STATE; }
    return sp_decompress_p(d->sp, src, len, dst, data_pnt, significant);
}

void ora_stop_and_clean(ora_dec *d) { if (d->sp) sp_stop(d->sp); }

/* Pixels the codec never writes (width/height remainders mod 4 for MSVideo1) are carried over from
 * the previous picture so that every materialised frame is fully defined. */
static void carry_remainder(int X, int Y, int32_t *cur, const int32_t *prev)
{
    int bw = X & ~3, bh = Y & ~3;
    if (bw == X && bh == Y) return;
    for (int y = 0; y < Y; y++) {
        int x0 = (y < bh) ? bw : 0;
        for (int x = x0; x < X; x++) cur[(size_t)y * X + x] = prev ? prev[(size_t)y * X + x] : 0;
    }
}

static void decode_stream_impl(ora_dec *d, int n_frames, const uint8_t *bytes, const uint64_t *frame_off,
                               const uint32_t *frame_len, const uint8_t *frame_key,
                               int32_t *out, int32_t *ring[2],
                               uint8_t *changed, uint8_t *significant, int32_t *status)
{
    const size_t npix = (size_t)d->X * d->Y;
    const int32_t *prev_out = NULL;     /* last materialised picture */
    int ring_i = 0;
    for (int f = 0; f < n_frames; f++) {
        const uint8_t *src = bytes + frame_off[f];
        int len = (int)frame_len[f];
        int32_t *dst;
        if (out) dst = out + (size_t)f * npix;
        else { dst = ring[ring_i]; if (dst == ora_previous_frame(d)) { ring_i ^= 1; dst = ring[ring_i]; } }
        if (d->msv1) carry_remainder(d->X, d->Y, dst, prev_out);
        const int32_t *pnt = NULL; int sig = 0, st = ORA_ZERO_STATE;
        if (frame_key[f]) {
            st = ora_decompress_i(d, src, len, dst);
            pnt = ora_previous_frame(d);
        } else {
            st = ora_decompress_p(d, src, len, dst, &pnt, &sig);
        }
        int ch = (pnt == dst);
        if (out && !ch) {               /* unchanged: materialise a replica of the shown picture */
            if (pnt) memcpy(dst, pnt, npix * 4); else memset(dst, 0, npix * 4);
        }
        if (changed) changed[f] = (uint8_t)ch;
        if (significant) significant[f] = (uint8_t)sig;
        if (status) status[f] = st;
        prev_out = dst;
        if (!out) ring_i ^= 1;
    }
}

int ora_decode_stream(int codec, int width, int height, int bpp, const uint8_t *palette, int palette_bytes,
                      int insignificant_lines, int n_frames, const uint8_t *bytes, const uint64_t *frame_off,
                      const uint32_t *frame_len, const uint8_t *frame_key,
                      int32_t *out, uint8_t *changed, uint8_t *significant, int32_t *status);

/* ---- the caller's side of the path (reference src/Manager.hx), restated for the "next" rows of SURVEY.md 8f ---- */

/* Manager.fill_bitmap_data, canvas branch (Manager.hx:363-381): 0x00RRGGBB -> the Int32 view of canvas bytes R,G,B,A
 * with alpha 255; ScreenPressor at 16 bpp stores 5-bit channels and is shifted instead (convert_fromRGB15,
 * Manager.hx:120,366-370).  flip: the vertical flip Main applies at render time (Main.hx:318,946). */
void ora_display_convert(const int32_t *src, int32_t *dst, int X, int Y, int from_rgb15, int flip)
{
    for (int y = 0; y < Y; y++) {
        const int32_t *s = src + (size_t)y * X;
        int32_t *d = dst + (size_t)(flip ? Y - 1 - y : y) * X;
        for (int x = 0; x < X; x++) {
            const uint32_t c = (uint32_t)s[x];
            d[x] = (int32_t)(from_rgb15 ? (0xFF000000u | (c << 3))
                                        : (0xFF000000u | ((c & 0xFF) << 16) | (c & 0xFF00) | ((c >> 16) & 0xFF)));
        }
    }
}

/* Manager.frames_differ_significantly (Manager.hx:392-421) for key frame n of a stream: prev_key / prev_data = the
 * previous frame record, pnt1 = the new picture, pnt2 = the picture shown before it (NULL: nothing yet -> the
 * reference would throw; defined as "differs"). */
int ora_frames_differ(int n, int prev_key, const uint8_t *prev_data, int prev_len, const uint8_t *cur_data, int cur_len,
                      const int32_t *pnt1, const int32_t *pnt2, int X, int Y, int insignificant_lines)
{
    if (n > 0) {
        if (prev_key && prev_data) {
            if (prev_len == cur_len) return memcmp(prev_data, cur_data, (size_t)cur_len) != 0;
            return 1;
        }
    } else return 1;
    if (!pnt2) return 1;
    for (size_t i = (size_t)insignificant_lines * X; i < (size_t)X * Y; i++)
        if (pnt1[i] != pnt2[i]) return 1;
    return 0;
}

/* ora_decode_stream + differs[f] = Manager's significance of key frame f (0 for non-key frames) */
int ora_decode_stream_differs(int codec, int width, int height, int bpp, const uint8_t *palette, int palette_bytes,
                              int insignificant_lines, int n_frames, const uint8_t *bytes, const uint64_t *frame_off,
                              const uint32_t *frame_len, const uint8_t *frame_key, int32_t *out, uint8_t *differs)
{
    const size_t npix = (size_t)width * height;
    ora_decode_stream(codec, width, height, bpp, palette, palette_bytes, insignificant_lines, n_frames, bytes, frame_off,
                      frame_len, frame_key, out, NULL, NULL, NULL);
    for (int f = 0; f < n_frames; f++) {
        differs[f] = 0;
        if (!frame_key[f]) continue;
        differs[f] = (uint8_t)ora_frames_differ(f, f > 0 ? frame_key[f - 1] : 0, f > 0 ? bytes + frame_off[f - 1] : NULL,
                                                f > 0 ? (int)frame_len[f - 1] : 0, bytes + frame_off[f], (int)frame_len[f],
                                                out + (size_t)f * npix, f > 0 ? out + (size_t)(f - 1) * npix : NULL,
                                                width, height, insignificant_lines);
    }
    return 0;
}

int ora_decode_stream(int codec, int width, int height, int bpp, const uint8_t *palette, int palette_bytes,
                      int insignificant_lines, int n_frames, const uint8_t *bytes, const uint64_t *frame_off,
                      const uint32_t *frame_len, const uint8_t *frame_key,
                      int32_t *out, uint8_t *changed, uint8_t *significant, int32_t *status)
{
    ora_dec *d = ora_create(codec, width, height, bpp, palette, palette_bytes);
    ora_preinit(d, insignificant_lines);
    decode_stream_impl(d, n_frames, bytes, frame_off, frame_len, frame_key, out, NULL, changed, significant, status);
    ora_destroy(d);
    return 0;
}

/* ---- multi-threaded baseline driver ---- */
typedef struct {
    const ora_stream_desc *streams; int n_streams; int insign;
    volatile int *next; pthread_mutex_t *mu;
} mt_ctx;

static void *mt_worker(void *arg)
{
    mt_ctx *c = (mt_ctx *)arg;
    int32_t *ring[2] = { NULL, NULL }; size_t ring_px = 0;
    for (;;) {
        pthread_mutex_lock(c->mu);
        int s = *c->next; if (s < c->n_streams) *c->next = s + 1;
        pthread_mutex_unlock(c->mu);
        if (s >= c->n_streams) break;
        const ora_stream_desc *sd = &c->streams[s];
        size_t npix = (size_t)sd->width * sd->height;
        if (!sd->out && npix > ring_px) {
            free(ring[0]); free(ring[1]);
            ring[0] = (int32_t *)calloc(npix, 4); ring[1] = (int32_t *)calloc(npix, 4); ring_px = npix;
        }
        ora_dec *d = ora_create(sd->codec, sd->width, sd->height, sd->bpp, sd->palette, sd->palette_bytes);
        ora_preinit(d, c->insign);
        decode_stream_impl(d, sd->n_frames, sd->bytes, sd->frame_off, sd->frame_len, sd->frame_key,
                           sd->out, ring, NULL, NULL, NULL);
        ora_destroy(d);
    }
    free(ring[0]); free(ring[1]);
    return NULL;
}

double ora_decode_streams_mt(const ora_stream_desc *streams, int n_streams, int n_threads,
                             int insignificant_lines, uint64_t *pixels)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_streams) n_threads = n_streams > 0 ? n_streams : 1;
    pthread_t *th = (pthread_t *)calloc(n_threads, sizeof *th);
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    volatile int next = 0;
    mt_ctx c = { streams, n_streams, insignificant_lines, &next, &mu };
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, mt_worker, &c);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    uint64_t px = 0;
    for (int s = 0; s < n_streams; s++) px += (uint64_t)streams[s].width * streams[s].height * streams[s].n_frames;
    if (pixels) *pixels = px;
    return (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
}
