// sp_common.cuh -- ScreenPressor frame decode for sm_100a, shared by the range-coder (v2) and rANS (v3/v4)
// entropy kernels.  Replaces reference src/ScreenPressor.hx:117-295 (DecompressI) and :302-484 (DecompressP).
//
// One warp decodes one frame of one stream: the entropy decode is a strictly serial chain through the coder
// state and the adaptive models (every symbol depends on the previous one, and the context of the next colour
// depends on reconstructed pixels -- ScreenPressor.hx:274-275,462-463), so the parallelism is (a) across
// streams: one warp-sized CTA per stream, hundreds per launch, and (b) inside a symbol: the 32 lanes search the
// cumulative-frequency table together and write / predict a run of pixels together.  Whole-picture work
// (P-frame "start from the previous picture", unchanged frames, flat frames) is done by the wide frame_copy
// kernel before this one (frame_ops.cu).
#pragma once
#include "common.cuh"

namespace jsp {

enum : uint32_t {
    SPJ_IFRAME   = 1u << 0,   // coded I frame (head & 15 == 2)
    SPJ_RENEW    = 1u << 1,   // flat I frame: only reset the models (RenewI, ScreenPressor.hx:108-115)
    SPJ_DIFF16   = 1u << 2,   // 16 bpp stream on the range coder: other context constants (:200-202, :316-318)
    SPJ_CXSHIFT0 = 1u << 3,   // SC_CXSHIFT == 0 (:59)
    SPJ_ANS      = 1u << 5,   // rANS stream (v3 / v4): sp_ans.cuh, else the range coder of sp_rc.cuh
    SPJ_ANS_V3   = 1u << 4,   // rANS stream version 3: Cx6.f0 = 64 (v4: 32), ScreenPressor.hx:69-72
};

struct SpJob {
    const uint8_t *src;       // compressed frame (device)
    int32_t       *dst;       // output picture; for P frames it already holds a copy of the previous picture
    const int32_t *prev;      // previous picture (P frames)
    uint32_t      *status;
    void          *state;     // per-stream model state (RcState / AnsState)
    uint8_t       *bts;       // per-stream block-type scratch, nbx*nby bytes
    uint32_t len, X, Y, flags;
    uint32_t insign_blocks;   // nbx * ceil(insignificant_lines / 16) (ScreenPressor.hx:86-89)
    uint32_t pad;
    uint32_t      *symbols;   // entropy-coded symbols this frame decoded (reporting: symbols / s), may be null
    uint32_t      *done;      // set to 1 (host-mapped memory) when the picture is complete and visible system-wide; may be null
};

constexpr uint32_t FULLMASK = 0xffffffffu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// per-byte add / subtract without carries between bytes (the predictor works per channel, mod 256)
__device__ __forceinline__ uint32_t vadd4(uint32_t a, uint32_t b) { return ((a & 0x7F7F7F7Fu) + (b & 0x7F7F7F7Fu)) ^ ((a ^ b) & 0x80808080u); }
__device__ __forceinline__ uint32_t vsub4(uint32_t a, uint32_t b) { return ((a | 0x80808080u) - (b & 0x7F7F7F7Fu)) ^ ((a ^ ~b) & 0x80808080u); }

__device__ __forceinline__ uint32_t px_load(const int32_t *p, long i, long end)
{
    return (p && i >= 0 && i < end) ? (uint32_t)p[i] : 0u;      // out-of-frame reads yield 0 (JavaScript undefined -> 0)
}

// One segment of m <= 32 consecutive pixels starting at index i, all produced by predictor `ptype`
// (ScreenPressor.hx:242-273 / :439-450).  `left` = the pixel at i-1.  Returns the segment's last pixel.
__device__ __forceinline__ uint32_t sp_segment(int32_t *dst, const int32_t *prev, long i, int m, int ptype,
                                               uint32_t clr, uint32_t left, long X, long end)
{
    const int lane = (int)lane_id();
    const long idx = i + lane;
    uint32_t v = clr;
    switch (ptype) {
    case 1: v = left; break;
    case 2: v = px_load(dst, idx - X, end); break;
    case 3: v = px_load(prev, idx, end); break;
    case 5: v = px_load(dst, idx - X - 1, end); break;
    case 4: {
        // JavaScript: a neighbour at a negative index is `undefined`, the byte sum is NaN and NaN & 0xFF is 0 -- the WHOLE
        // pixel is 0 when its above-left neighbour (the lowest index of the three) is outside (ScreenPressor.hx:443-449).
        // Such pixels (index <= X) are a prefix of the segment; the chain restarts from 0 after them.
        uint32_t d = (lane < m && idx > X) ? vsub4(px_load(dst, idx - X, end), px_load(dst, idx - X - 1, end)) : 0u;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t o = __shfl_up_sync(FULLMASK, d, s);
            if (lane >= s) d = vadd4(d, o);
        }
        v = vadd4(i > X ? left : 0u, d) & 0x00FFFFFFu;
        break;
    }
    default: break;
    }
    if (lane < m && idx >= 0 && idx < end) dst[idx] = (int32_t)v;
    const uint32_t last = __shfl_sync(FULLMASK, v, (m - 1) & 31);
    __syncwarp();
    return last;
}

// I frames: predictors 2 / 4 / 5 read the row above (ScreenPressor.hx:253-272).  Reading it back from global memory would
// put an L2 round trip on the serial chain of every such run (stores do not stay in L1), so the last X + 1 pixels also
// live in a shared-memory ring ("a ring of the last X+1 pixels is sufficient").
__device__ __forceinline__ uint32_t sp_ring_size(uint32_t X) { return 1u << (32 - __clz(X + 65u)); }   // power of two > X + 65

__device__ __forceinline__ uint32_t sp_segment_ring(int32_t *dst, uint32_t *ring, uint32_t rmask, long i, int m, int ptype,
                                                    uint32_t clr, uint32_t left, long X, long end)
{
    const int lane = (int)lane_id();
    const long idx = i + lane;
    auto rd = [&](long j) -> uint32_t { return j >= 0 ? ring[(uint32_t)j & rmask] : 0u; };
    uint32_t v = clr;
    switch (ptype) {
    case 1: v = left; break;
    case 2: v = rd(idx - X); break;
    case 5: v = rd(idx - X - 1); break;
    case 4: {
        uint32_t d = lane < m ? vsub4(rd(idx - X), rd(idx - X - 1)) : 0u;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t o = __shfl_up_sync(FULLMASK, d, s);
            if (lane >= s) d = vadd4(d, o);
        }
        v = vadd4(left, d) & 0x00FFFFFFu;
        break;
    }
    default: break;
    }
    if (ptype > 1) __syncwarp();                    // every lane has read the row above before its slots are reused
    if (lane < m && idx < end) { dst[idx] = (int32_t)v; ring[(uint32_t)idx & rmask] = v; }
    // predictors 0 / 1 write one value everywhere: no broadcast needed for "the run's last pixel"
    const uint32_t last = ptype > 1 ? __shfl_sync(FULLMASK, v, (m - 1) & 31) : v;
    __syncwarp();
    return last;
}

// The host may copy a picture out while the launch that produced it is still running (other warps of the launch decode
// longer frames): every lane makes its stores visible system-wide, then one lane raises the frame's flag in mapped memory.
__device__ __forceinline__ void sp_signal_done(const SpJob &J)
{
    if (!J.done) return;
    __threadfence_system();
    __syncwarp();
    if (lane_id() == 0) *reinterpret_cast<volatile uint32_t *>(J.done) = 1u;
}

// A frame whose entropy decode failed shows the previous picture (a P frame returns the retained buffer) or
// nothing (a failed I frame has already dropped prevFrame, ScreenPressor.hx:110): undo the partial writes.
__device__ __forceinline__ void sp_undo_frame(const SpJob &J, bool iframe)
{
    const size_t n = (size_t)J.X * J.Y;
    const int lane = (int)lane_id();
    const int32_t *src = iframe ? nullptr : J.prev;
    __syncwarp();
    for (size_t i = lane; i < n; i += 32) J.dst[i] = src ? src[i] : 0;
    __syncwarp();
}

// cx + cx1 stays below 4096 on every valid stream (cx < 64, cx1 <= 0xFC0; 16 bpp v2: 5-bit channel values).  A corrupt
// stream can exceed it -- the reference would read outside cntab[] -- so the index wraps inside its channel and the
// frame is reported as failed; never an out-of-bounds access on the device.
template <class Coder>
__device__ __forceinline__ int sp_ctx_index(Coder &ec, int channel, int cx, int cx1)
{
    const int i = cx + cx1;
    if ((unsigned)i >= 4096u) ec.fail_frame();     // (two selects, no branch: the index is on every colour symbol's chain)
    return channel * 4096 + (i & 4095);
}

// Zero-length runs are legal syntax and cost almost no bits once their model has adapted, so a hostile stream could keep
// a warp busy for ever.  No encoder emits them: a frame that needs more run-loop iterations than this is failed.
__device__ __forceinline__ long sp_run_budget(long X, long Y) { return 2 * X * Y + 16 * ((X + 15) / 16) * ((Y + 15) / 16) + 4096; }

// Optional section timing (build with JSP_NVCC_EXTRA=-DJSP_PROFILE_SECTIONS): cycles spent per warp in the pieces of the
// I-frame loop, summed into g_sp_prof[] -- how the per-symbol latency was broken down.
#ifdef JSP_PROFILE_SECTIONS
__device__ unsigned long long g_sp_prof[8];
__device__ unsigned long long g_spp_prof[10];     // P frames: block types, symbol decodes, run writes, motion, whole, runs, row pieces, frames
#define JSP_PT0 const long long _pt0 = clock64();
#define JSP_PT1(k) _pacc[k] += clock64() - _pt0;
#define JSP_PCOUNT(k) _pacc[k]++;
#define JSP_T0 const long long _t0 = clock64();
#define JSP_T1(k) _acc[k] += clock64() - _t0;
#define JSP_PROF_DECL long long _acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define JSP_PROF_FLUSH if (lane_id() == 0) { for (int _k = 0; _k < 8; _k++) atomicAdd(&g_sp_prof[_k], (unsigned long long)_acc[_k]); }
#else
#define JSP_T0
#define JSP_T1(k)
#define JSP_PROF_DECL
#define JSP_PROF_FLUSH
#define JSP_PT0
#define JSP_PT1(k)
#define JSP_PCOUNT(k)
#endif

// ---- the frame loops, generic over the entropy coder (EntroCoder interface, EntroCoders.hx:8-24) ----
template <class Coder>
__device__ void sp_decode_iframe(Coder &ec, const SpJob &J, uint32_t *ring)
{
    const uint32_t rmask = sp_ring_size(J.X) - 1u;
    const long X = J.X, end = (long)J.X * J.Y;
    int32_t *dst = J.dst;
    const int cxshift = (J.flags & SPJ_CXSHIFT0) ? 0 : 2;
    int maskcx1 = 0xFC00, shiftcx1 = 4, shiftcx = 18;
    if (J.flags & SPJ_DIFF16) { maskcx1 = 0xFF00; shiftcx1 = 2; shiftcx = 16; }
    const int lane = (int)lane_id();
    const int chunk = X < 32 ? (int)X : 32;       // a chunk never reads pixels it writes itself (row above is X away)
    ec.renewI();
    ec.decodeBegin(J.src, J.len, 1);
    int cx = 0, cx1 = 0;
    long di = 0, k = 0;
    uint32_t clr = 0, lastval = 0;
    // ScreenPressor.hx:173-183.  A rolled loop on purpose: one copy of the colour decoder per call site instead of three
    // keeps the hot loop's code small (several warps share an SM's instruction cache).
    auto decode_rgb = [&]() -> uint32_t {
        uint32_t px = 0;
#pragma unroll 1
        for (int ch = 0; ch < 3; ch++) {
            const int v = ec.decodeClr(sp_ctx_index(ec, ch, cx, cx1));
            cx1 = (cx << 6) & 0xFC0; cx = v >> cxshift;
            px += (uint32_t)v << (8 * ch);
        }
        return px;
    };
    long budget = sp_run_budget(X, J.Y);
    while (k < X + 1) {                            // first X+1 pixels: (colour, run) pairs, :170-197
        if (--budget < 0) ec.fail_frame();
        clr = decode_rgb();
        const int n = ec.decodeN(0);
        if (ec.failed()) return;
        k += n;
        for (int o = 0; o < n; o += 32) {
            const long idx = di + o + lane;
            if (o + lane < n && idx < end) { dst[idx] = (int32_t)clr; if (ring) ring[(uint32_t)idx & rmask] = clr; }
        }
        if (n > 0) lastval = clr;
        di += n;
    }
    __syncwarp();
    int ptype = 0;
    JSP_PROF_DECL
#ifdef JSP_PROFILE_SECTIONS
    const long long _tall = clock64();
#endif
    while (di < end) {                             // :218-286
        if (--budget < 0) ec.fail_frame();
        { JSP_T0 ptype = ec.decodeP(ptype); JSP_T1(0) }
        if (ptype == 0) { JSP_T0 clr = decode_rgb(); JSP_T1(1) }
        int n;
        { JSP_T0 n = ec.decodeN(ptype); JSP_T1(2) }
        if (ec.failed()) return;
        if (ptype == 3 || ptype > 5) n = 0;        // no such predictor in an I frame: nothing is written
        if (ptype == 1) clr = lastval;             // `clr = dst[lasti]` even for an empty run (:252)
        { JSP_T0
        for (int o = 0; o < n; o += chunk) {
            const int m = n - o < chunk ? n - o : chunk;
            lastval = ring ? sp_segment_ring(dst, ring, rmask, di + o, m, ptype, clr, lastval, X, end)
                           : sp_segment(dst, nullptr, di + o, m, ptype, clr, lastval, X, end);
        }
        JSP_T1(3) }
        if (n > 0 && ptype != 0) clr = lastval;
        di += n;
        cx1 = ((int)clr & maskcx1) >> shiftcx1;    // :274-275
        cx = (int)clr >> shiftcx;
#ifdef JSP_PROFILE_SECTIONS
        _acc[5]++;
#endif
    }
#ifdef JSP_PROFILE_SECTIONS
    _acc[4] += clock64() - _tall;
#endif
    JSP_PROF_FLUSH
}

// A row piece decoded one pixel at a time by every lane (each reads back its own stores): for pieces that leave the
// picture, where pixels are not written and read as 0.  Cold: only corrupt streams get here.
static __device__ __noinline__ uint32_t sp_piece_pixelwise(int32_t *dst, const int32_t *prev, long i, int m, int ptype, uint32_t clr, long X, long end)
{
    uint32_t v = clr;
    for (int k = 0; k < m; k++) {
        const long p = i + k;
        switch (ptype) {
        case 1: v = px_load(dst, p - 1, end); break;
        case 2: v = px_load(dst, p - X, end); break;
        case 3: v = px_load(prev, p, end); break;
        case 4: v = p > X ? vadd4(px_load(dst, p - 1, end), vsub4(px_load(dst, p - X, end), px_load(dst, p - X - 1, end))) & 0x00FFFFFFu : 0u; break;
        case 5: v = px_load(dst, p - X - 1, end); break;
        default: break;
        }
        if (p >= 0 && p < end) dst[p] = (int32_t)v;
    }
    return v;
}

// P-frame data blocks are decoded against a shared-memory copy of the rectangle and its upper / left neighbours: every
// row piece of a run reads the pixel to its left or the row above, and from global memory each of those reads is an L2
// round trip on the serial chain (~550 cycles per row piece measured, tools/sp_pframe_profile.py).
constexpr int SP_PT_STRIDE = 17;                                   // (16 + 1) x (16 + 1) pixels of the output picture ...
constexpr int SP_PT_PREV = 320;                                   // ... then 16 x 16 of the previous picture
constexpr uint32_t SP_PTILE_WORDS = SP_PT_PREV + 16 * 16 + 32;     // + slack for the reads of inactive lanes

template <class Coder>
__device__ void sp_decode_pframe(Coder &ec, const SpJob &J, uint32_t &status_bits, uint32_t *ptile)
{
    const long X = J.X, Y = J.Y, end = X * Y;
    const int nbx = (int)((J.X + 15) / 16), nby = (int)((J.Y + 15) / 16), nb = nbx * nby;
    int32_t *dst = J.dst;
    const int32_t *prev = J.prev;
    uint8_t *bts = J.bts;
    const int lane = (int)lane_id();
    const int cxshift = (J.flags & SPJ_CXSHIFT0) ? 0 : 2;
    int maskcx1 = 0xFC00, shiftcx1 = 4, shiftcx = 18;
    if (J.flags & SPJ_DIFF16) { maskcx1 = 0xFF00; shiftcx1 = 2; shiftcx = 16; }
#ifdef JSP_PROFILE_SECTIONS
    long long _pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long _pall = clock64();
    struct PFlush { long long *a; long long t0; __device__ ~PFlush() { a[4] += clock64() - t0; a[7]++;
        if (lane_id() == 0) { for (int k = 0; k < 8; k++) atomicAdd(&g_spp_prof[k], (unsigned long long)a[k]);
                              atomicMax(&g_spp_prof[8], (unsigned long long)a[4]); } } } _pflush{_pacc, _pall};
#endif
    ec.decodeBegin(J.src, J.len, 1);
    uint32_t xw = 0;                               // four decodeX symbols: xx1 = lo | hi << 8, xx2 likewise (:322-327); one call site
#pragma unroll 1
    for (int i = 0; i < 4; i++) xw |= (uint32_t)(ec.decodeX() & 0xFF) << (8 * i);
    const int xx1 = (int)(xw & 0xFFFFu), xx2 = (int)(xw >> 16);
    for (int i = lane; i < nb; i += 32) bts[i] = 0;
    __syncwarp();
    bool signif = false;
    long x = xx1;
    long budget = sp_run_budget(X, Y);
    while (x <= xx2) {                             // block types, run-length coded (:336-344)
        if (--budget < 0) ec.fail_frame();
        const int bt = ec.decodeBT();
        const int n = ec.decodeBN();
        if (ec.failed()) return;
        for (int o = lane; o < n; o += 32) { const long p = x + o; if (p >= 0 && p < nb) bts[p] = (uint8_t)bt; }
        if (bt > 0 && n > 0) {                     // significant changes? (:347-352)
            const long lo = x > (long)J.insign_blocks ? x : (long)J.insign_blocks;
            const long hi = x + n - 1 < nb - 1 ? x + n - 1 : nb - 1;
            if (lo <= hi) signif = true;
        }
        x += n;
    }
    __syncwarp();
#ifdef JSP_PROFILE_SECTIONS
    _pacc[0] += clock64() - _pall;
#endif
    if (signif) status_bits |= ST_SIGNIFICANT;
    status_bits |= ST_CHANGED;
    int cx = 0, cx1 = 0;
    uint32_t clr = 0;
    int lastmx = 0, lastmy = 0;
    auto decode_rgb = [&]() -> uint32_t {
        uint32_t px = 0;
#pragma unroll 1
        for (int ch = 0; ch < 3; ch++) {
            const int v = ec.decodeClr(sp_ctx_index(ec, ch, cx, cx1));
            cx1 = (cx << 6) & 0xFC0; cx = v >> cxshift;
            px += (uint32_t)v << (8 * ch);
        }
        return px;
    };
    for (int bi = 0; bi < nb; bi++) {
        // skip unchanged blocks 32 at a time: they already hold the previous picture
        if ((bi & 31) == 0) {
            const int probe = bi + lane < nb ? bts[bi + lane] : 0;
            const uint32_t any = __ballot_sync(FULLMASK, probe != 0);
            if (!any) { bi += 31; continue; }
        }
        const int bt = bts[bi];
        if (bt == 0) continue;
        const int by = bi / nbx, bx = bi - by * nbx;
        const int y16 = by * 16, x16 = bx * 16;
        int x1 = x16, x2 = x16 + 16, y1 = y16, y2 = y16 + 16;
        if (x2 > X) x2 = (int)X;
        if (y2 > Y) y2 = (int)Y;
        JSP_PT0
        if (((bt - 1) & 1) > 0) {                  // sub-rectangle (:375-386); the block itself is already copied
            uint32_t sw = 0;                       // four decodeSXY symbols (0..15 each); one call site
#pragma unroll 1
            for (int i = 0; i < 4; i++) sw |= (uint32_t)(ec.decodeSXY(i) & 0xFF) << (8 * i);
            x1 = (int)(sw & 0xFFu) + x16;
            y1 = (int)((sw >> 8) & 0xFFu) + y16;
            x2 = (int)((sw >> 16) & 0xFFu) + x16 + 1;
            y2 = (int)(sw >> 24) + y16 + 1;
        }
        if (((bt - 1) & 2) > 0) {                  // motion vector (:388-405)
            int mx, my;
            if (Coder::kCanDecodeBool && ec.decodeBool()) { mx = lastmx; my = lastmy; }
            else { mx = ec.decodeMX() - 256; my = ec.decodeMY() - 256; }
            if (ec.failed()) return;
            lastmx = mx; lastmy = my;
            const int w = x2 - x1;                 // <= 16: the rectangle lies inside one block
            for (int y = y1; y < y2; y += 2) {     // two rows of up to 16 pixels per pass
                const int ry = y + (lane >> 4), rx = lane & 15;
                if (ry < y2 && rx < w) {
                    const long i = (long)ry * X + x1 + rx, j = (long)(ry + my) * X + (x1 + mx) + rx;
                    if (i >= 0 && i < end) dst[i] = (int32_t)px_load(prev, j, end);
                }
            }
            __syncwarp();
            JSP_PT1(3)
        } else {                                   // data (:406-467): runs in raster order inside the rectangle
            JSP_PT1(3)
            int xq = x1, y = y1, ptype = 0;
            // well-formed rectangle: stage it (rows y1-1 .. y2-1, columns x1-1 .. x2-1, by linear pixel index) in shared memory
            // (not when the rectangle spans the picture's width: the pixel "left" of column 0 is the previous row's last
            // pixel, which the block itself rewrites)
            const bool tiled = ptile != nullptr && x2 > x1 && x2 <= X && y2 > y1 && y2 <= Y && !(x1 == 0 && x2 == X);
            if (tiled) {
                const int w = x2 - x1, h = y2 - y1;
                // all loads first, then the stores: the warp waits for one L2 round trip, not for eighteen
                uint32_t ra[10], rb[8];
#pragma unroll
                for (int it = 0; it < 10; it++) {
                    const int k = lane + 32 * it, ry = k / 17, rx = k - ry * 17;
                    ra[it] = (ry <= h && rx <= w) ? px_load(dst, (long)(y1 - 1 + ry) * X + (x1 - 1 + rx), end) : 0u;
                }
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int k = lane + 32 * it, ry = k >> 4, rx = k & 15;
                    rb[it] = (ry < h && rx < w) ? px_load(prev, (long)(y1 + ry) * X + (x1 + rx), end) : 0u;
                }
#pragma unroll
                for (int it = 0; it < 10; it++) ptile[lane + 32 * it] = ra[it];       // 320 words: up to SP_PT_PREV
#pragma unroll
                for (int it = 0; it < 8; it++) ptile[SP_PT_PREV + lane + 32 * it] = rb[it];
                __syncwarp();
            }
            while (y < y2) {
                if (--budget < 0) ec.fail_frame();
                int n;
                { JSP_PT0
                ptype = ec.decodeP(ptype);
                if (ptype == 0) clr = decode_rgb();
                n = ec.decodeN(ptype);
                JSP_PT1(1) JSP_PCOUNT(5) }
                if (ec.failed()) return;
                JSP_PT0
                while (n > 0) {
                    JSP_PCOUNT(6)                    // one row piece at a time: later rows may read this one
                    int m = x2 - xq; if (m > n) m = n; if (m > 32) m = 32;
                    if (m <= 0) {                  // degenerate rectangle (x2 <= x1): the reference steps one pixel per row
                        const long i = (long)y * X + xq;
                        uint32_t v = clr;
                        if (ptype == 1) v = px_load(dst, i - 1, end); else if (ptype == 2) v = px_load(dst, i - X, end);
                        else if (ptype == 3) v = px_load(prev, i, end); else if (ptype == 5) v = px_load(dst, i - X - 1, end);
                        else if (ptype == 4) v = i > X ? vadd4(px_load(dst, i - 1, end), vsub4(px_load(dst, i - X, end), px_load(dst, i - X - 1, end))) & 0xFFFFFFu : 0u;
                        if (lane == 0 && i >= 0 && i < end) dst[i] = (int32_t)v;
                        __syncwarp();
                        clr = v; n--; xq = x1; y++;
                        continue;
                    }
                    const long i = (long)y * X + xq;
                    uint32_t last;
                    if (tiled && y < y2) {
                        uint32_t *row = ptile + (y - y1 + 1) * SP_PT_STRIDE + (xq - x1 + 1);      // &tile[y][xq]
                        uint32_t v = clr;
                        switch (ptype) {
                        case 1: v = row[-1]; break;
                        case 2: v = row[lane - SP_PT_STRIDE]; break;
                        case 3: v = ptile[SP_PT_PREV + (y - y1) * 16 + (xq - x1) + lane]; break;
                        case 5: v = row[lane - SP_PT_STRIDE - 1]; break;
                        case 4: {
                            // pixels whose above-left neighbour has a negative index are 0 (see sp_segment)
                            uint32_t d = (lane < m && i + lane > X) ? vsub4(row[lane - SP_PT_STRIDE], row[lane - SP_PT_STRIDE - 1]) : 0u;
#pragma unroll
                            for (int sft = 1; sft < 16; sft <<= 1) {           // m <= 16 here
                                const uint32_t o = __shfl_up_sync(FULLMASK, d, sft);
                                if (lane >= sft) d = vadd4(d, o);
                            }
                            v = vadd4(i > X ? row[-1] : 0u, d) & 0x00FFFFFFu;
                            break;
                        }
                        default: break;
                        }
                        if (lane < m) { row[lane] = v; dst[i + lane] = (int32_t)v; }
                        last = __shfl_sync(FULLMASK, v, (m - 1) & 31);
                        __syncwarp();
                    } else if (i + m > end || m >= X) {
                        // a piece that leaves the picture (or is wider than a picture row): only a corrupt stream gets
                        // here.  Pixels outside the picture are not written and read as 0, so the left-to-right chain
                        // the parallel segment assumes does not hold: one pixel at a time, as the reference does it
                        // (every lane runs the same loop and reads back its own stores)
                        last = sp_piece_pixelwise(dst, prev, i, m, ptype, clr, X, end);
                        __syncwarp();
                    } else {
                        const uint32_t left = (ptype == 1 || ptype == 4) ? px_load(dst, i - 1, end) : 0u;
                        last = sp_segment(dst, prev, i, m, ptype, clr, left, X, end);
                    }
                    if (ptype != 0) clr = last;
                    n -= m; xq += m;
                    if (xq >= x2) { xq = x1; y++; }
                }
                JSP_PT1(2)
                cx1 = ((int)clr & maskcx1) >> shiftcx1;    // :462-463
                cx = (int)clr >> shiftcx;
            }
        }
    }
}

}  // namespace jsp
