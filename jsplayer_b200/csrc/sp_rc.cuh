// sp_rc.cuh -- ScreenPressor v2 entropy decode on sm_100a: 32-bit byte-wise range decoder with adaptive
// frequency tables.  Replaces reference src/RangeCoder.hx (whole file) and EntroCoderRC
// (src/EntroCoders.hx:31-180) for one stream per warp.
//
//  * coder state (range, code, read position) is replicated in every lane: no broadcasts on the serial chain;
//  * the non-colour tables (ntab 6x257, ptypetab 6x7, xxtab, ntab2, bttab, sxytab 4x17, mvtab 2x513 -- 15 KB in
//    the prefix layout below) live in shared memory for the whole frame and are saved to / restored from the
//    stream's state in HBM; the bitstream is read through a 128-byte shared-memory window;
//  * a colour context is one 1280-byte row in HBM (lane-local prefix sums of 256 counts + 32 lane bases + total +
//    generation tag; the reference's 16 group sums are derived data and are not stored).  Rows are DECODED in a
//    12-slot LRU cache in shared memory, exactly like the small tables: a global store invalidates the L1 line it
//    hits and every symbol updates its row, so decoding rows in place would pay an L2 round trip per symbol.
//    The symbol search is ONE __ballot_sync plus a 3-probe binary search in registers -- no prefix scan, no linear
//    search (RangeCoder.hx:58-65, :90-108) -- and a single division per symbol (see "table layout");
//  * renewI (EntroCoders.hx:81-130) is O(1) for the 12288 colour rows: it bumps a generation number and rows
//    with an older tag read as "all ones" (the reference also resets lazily, :85).
#pragma once
#include "sp_common.cuh"
#include <cstddef>
#include <cstring>

namespace jsp {

constexpr uint32_t RC_TOP = 0x01000000u, RC_BOT = 0x010000u;   // RangeCoder.hx:12-13

// ---- table layout ------------------------------------------------------------------------------------------
// The reference keeps raw counts and finds a symbol by a linear cumulative search (RangeCoder.hx:58-65, :90-108).
// Here a table of 32*K symbols stores, per lane, the INCLUSIVE prefix sums of the lane's own K counts (lp) and
// the lane's exclusive base (sum of all counts of lower lanes).  Decoding needs no prefix scan at all: a lane
// compares (base + lp[q]) * r with `code`, one __ballot_sync finds the lane, and an update adds `step` to at
// most K values in one lane plus the bases of the higher lanes.  Counts are recovered as differences only when
// the table is rescaled (total > BOT).  The symbol found is the reference's: the search is a pure function of
// the cumulative counts, and cum <= code / r  <=>  cum * r <= code for integers -- which also removes the
// reference's second division (get_freq, RangeCoder.hx:45-49) from the serial chain.
template <int K>
struct RcBig {                                                 // K = 8: 256 symbols, K = 16: 512 symbols
    uint32_t lp[32 * K];
    uint32_t base[32];
    uint32_t total;
    uint32_t tag;                                              // colour rows: generation; shared-memory tables: unused
    uint32_t pad[2];
};
struct RcTiny {                                                // up to 32 symbols: inclusive cumulative counts, one per lane
    uint32_t P[32];
};
constexpr int RC_ROW_STRIDE = 320;                             // u32 per colour row in HBM (RcBig<8> = 292, padded to 1280 B)
constexpr int RC_ROWS = 3 * 4096;

struct RcSmall {
    RcBig<8> ntab[6], xxtab, ntab2;
    RcBig<16> mvtab[2];
    RcTiny sxytab[4], ptypetab[6], bttab;
    alignas(16) uint8_t win[128];                              // bitstream window (not state; lives here to share the allocation)
};

// Shared memory of one warp: the small tables (saved to / restored from RcState) and a cache of colour rows.
// A global store invalidates the L1 line it hits, and every symbol updates its row, so a row read straight from global
// memory pays an L2 round trip (~500 cycles) on EVERY symbol.  Screen content keeps returning to a handful of contexts
// (12 fully associative LRU slots catch 93-98 % of the accesses of the synthetic corpus), so rows are decoded in shared
// memory like the small tables and written back when evicted and at the end of the frame.
constexpr int RC_CACHE_ROWS = 12;
struct RcShared {
    RcSmall small;
    RcBig<8> cache[RC_CACHE_ROWS];
};

struct RcState {                                               // per stream, in HBM
    RcSmall small;
    uint32_t gen;                                              // generation of the colour rows (bumped by renewI)
    uint32_t pad[3];
    uint32_t *rows;                                            // RC_ROWS * RC_ROW_STRIDE u32, separately allocated
};

__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));    // one MUFU; the division below is exact by construction
    return y;
}

// floor(a / b) for 5 <= b <= 2^17 (a table total).  The float pipeline (I2F / F2I conversions, a correctly rounded
// reciprocal) costs ~260 cycles on the one warp that waits for it, so only the reciprocal of b -- which does not depend
// on the coder state -- is formed in floating point (one MUFU); it is turned into a 32-bit fixed-point reciprocal that
// is guaranteed not to exceed 2^32 / b, and the quotient comes from two multiply-highs plus exact corrections.
struct RcRecip { uint32_t inv; };
__device__ __forceinline__ RcRecip rc_recip(uint32_t b)
{
    // rb ~ 1 / b within 2^-22 relative; scale by 2^32 (b >= 5 keeps it below 2^30), then shave 2^-20 relative + 1 so that
    // inv <= floor(2^32 / b) whatever the rounding of the approximation was
    const uint32_t est = __float2uint_rz(rcp_approx(__uint2float_rn(b)) * 4294967296.0f);
    RcRecip r; r.inv = est - (est >> 20) - 1u;
    return r;
}
__device__ __forceinline__ uint32_t udiv_small(uint32_t a, uint32_t b, RcRecip rc)
{
    uint32_t q = __umulhi(a, rc.inv);                // <= a / b, short by at most a * 2^-19 / b + 1
    uint32_t rem = a - q * b;
    const uint32_t q2 = __umulhi(rem, rc.inv);       // the shortfall, again from below: now short by at most 2
    q += q2; rem -= q2 * b;
    if (rem >= b) { q++; rem -= b; }
    if (rem >= b) { q++; rem -= b; }
    while (rem >= b) { q++; rem -= b; }               // never runs; keeps the result exact by construction
    return q;
}

#ifdef JSP_PROFILE_SECTIONS
__device__ unsigned long long g_rc_prof[8];
#endif

struct RcCoder {
    static constexpr bool kCanDecodeBool = false;              // EntroCoders.hx:178
    RcSmall *sm;                                               // shared memory: small tables
    RcBig<8> *cache;                                           // shared memory: colour-row cache
    int my_tag;                                                // lane < RC_CACHE_ROWS: context index held by slot `lane`, -1 = empty
    uint32_t my_age, tick;                                     // LRU stamps
    uint32_t *rows;
    uint32_t gen;
    uint32_t range, code;
    const uint8_t *data;
    uint32_t len, pos, wbase;
    uint32_t nsym;                                             // symbols decoded in this frame (reporting only)
#ifdef JSP_PROFILE_SECTIONS
    long long prof[8];
#define JSP_PT(k) { const long long _n = clock64(); prof[k] += _n - _pt; _pt = _n; }
#define JSP_PT0 long long _pt = clock64();
#else
#define JSP_PT(k)
#define JSP_PT0
#endif
    bool poisoned, fail;

    __device__ __forceinline__ bool failed() const { return fail; }
    __device__ __forceinline__ void fail_frame() { fail = true; range = 0; }   // a failure found by the frame loop; see decode_tiny

    __device__ __forceinline__ void next_byte()
    {
        if (pos < len) {
            if (pos - wbase >= 128u) {                         // refill the window: one coalesced warp load
                __syncwarp();
                wbase = pos & ~127u;
                const int lane = (int)lane_id();
                uint32_t w = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t p = wbase + 4u * lane + k;
                    if (p < len) w |= (uint32_t)__ldg(data + p) << (8 * k);
                }
                reinterpret_cast<uint32_t *>(sm->win)[lane] = w;
                __syncwarp();
            }
            code = (code << 8) | sm->win[pos - wbase];
        } else poisoned = true;                                // JS: code becomes NaN for good (RangeCoder.hx:41)
        pos++;
    }
    __device__ void decodeBegin(const uint8_t *src, uint32_t n, uint32_t pos0)   // RangeCoder.hx:19-34
    {
        data = src; len = n; code = 0; range = 0xFFFFFFFFu; poisoned = false; wbase = 0x80000000u;
        pos = pos0 + 1;
        next_byte(); next_byte(); next_byte(); next_byte();
    }
    // RangeCoder.hx:36-43 with the products already formed: lo = cumFreq * r, width = freq * r
    __device__ __forceinline__ void consume(uint32_t lo, uint32_t width)
    {
        code -= lo;
        range = width;
        while (range < RC_TOP) { next_byte(); range <<= 8; }
    }

    template <int K>
    __device__ __forceinline__ void init_big(RcBig<K> &t)
    {
        const int lane = (int)lane_id();
#pragma unroll
        for (int q = 0; q < K; q++) t.lp[lane * K + q] = q + 1;
        t.base[lane] = K * lane;
        if (lane == 0) t.total = 32 * K;
    }
    __device__ __forceinline__ void init_tiny(RcTiny &t) { t.P[lane_id()] = lane_id() + 1; }

    __device__ void renewI()                                                     // EntroCoders.hx:81-130
    {
        gen = gen + 1;
        for (int t = 0; t < 6; t++) { init_big(sm->ntab[t]); init_tiny(sm->ptypetab[t]); }
        init_big(sm->xxtab); init_big(sm->ntab2);
        init_big(sm->mvtab[0]); init_big(sm->mvtab[1]);
        for (int t = 0; t < 4; t++) init_tiny(sm->sxytab[t]);
        init_tiny(sm->bttab);
        __syncwarp();
    }

    // RangeCoder.hx:51-80 for tables of N <= 32 symbols
    template <int N>
    __device__ int decode_tiny(RcTiny &t, uint32_t step)
    {
        const int lane = (int)lane_id();
        nsym++;
        uint32_t p = t.P[lane];
        uint32_t tot = __shfl_sync(FULLMASK, p, N - 1);
        // A failed frame decodes nothing more (the models stay as they were), at no cost to the symbol chain: failing
        // zeroes `range`, a zero range gives r = 0, every product is then 0 <= code and the search runs off the table
        // again.  A symbol asked for after the data ran out (`poisoned`) fails the same way.
        if (poisoned) range = 0;
        const uint32_t r = udiv_small(range, tot, rc_recip(tot));
        const uint32_t codev = code;
        const uint32_t pr = p * r;
        const int s = __popc(__ballot_sync(FULLMASK, lane < N && pr <= codev));
        if (s >= N) { range = 0; fail = true; return N - 1; }
        const uint32_t below = __shfl_sync(FULLMASK, pr, (s + 31) & 31), hi = __shfl_sync(FULLMASK, pr, s);
        const uint32_t lo = s ? below : 0u;
        consume(lo, hi - lo);
        if (lane >= s) p += step;
        tot += step;
        if (tot > RC_BOT) {                                    // :70-77: every count -> (count >> 1) + 1
            const uint32_t prev = __shfl_up_sync(FULLMASK, p, 1);
            uint32_t c = lane < N ? ((p - (lane ? prev : 0u)) >> 1) + 1 : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULLMASK, c, d); if (lane >= d) c += o; }
            p = c;
            t.P[lane] = p;
        } else if (lane >= s) t.P[lane] = p;
        return s;                                              // a lane reads back only its own P[lane]: no barrier needed
    }

    // RangeCoder.hx:51-80 (256 / 512 symbols) and :82-130 (colour rows, cached in shared memory; the reference's 16 group
    // sums are derived data and are not kept), on a table in shared memory
    template <int K>
    __device__ __forceinline__ int decode_big(uint32_t *tab, uint32_t step)
    {
        const int lane = (int)lane_id();
        nsym++;
        JSP_PT0
        uint32_t lp[K];
        uint32_t base, tot;
        {
            const uint4 *t4 = reinterpret_cast<const uint4 *>(tab) + lane * (K / 4);
#pragma unroll
            for (int q = 0; q < K / 4; q++) { const uint4 v = t4[q]; lp[4 * q] = v.x; lp[4 * q + 1] = v.y; lp[4 * q + 2] = v.z; lp[4 * q + 3] = v.w; }
            base = tab[32 * K + lane];
            tot = tab[32 * K + 32];
        }
        if (poisoned) range = 0;                               // as in decode_tiny: fails below, touches nothing
        JSP_PT(0)
        const uint32_t r = udiv_small(range, tot, rc_recip(tot));
        JSP_PT(1)
        const uint32_t codev = code;
        const uint32_t br = base * r;
        if (codev >= tot * r) { range = 0; fail = true; return 32 * K - 1; }     // value >= total: not a valid stream
        const int L = __popc(__ballot_sync(FULLMASK, br <= codev)) - 1;          // lane 0 has base 0: L >= 0
        const uint32_t t = codev - br;                                           // meaningful in lane L only
        // in lane L the symbol's inclusive prefix exceeds t (the next lane's base is above the value), so
        // m = #{q : lp[q] * r <= t} is at most K - 1: a binary search over lp[0 .. K-2] finds it together with
        // the two neighbouring products lo = lp[m-1] * r (0 if m = 0) and hi = lp[m] * r
        uint32_t lo = 0, hi = lp[K - 1] * r; int m = 0;
        if constexpr (K == 8) {
            const uint32_t pa = lp[3] * r; const bool a = pa <= t;
            if (a) lo = pa; else hi = pa;
            const uint32_t pb = (a ? lp[5] : lp[1]) * r; const bool bq = pb <= t;
            if (bq) lo = pb; else hi = pb;
            const uint32_t v = bq ? (a ? lp[6] : lp[2]) : (a ? lp[4] : lp[0]);
            const uint32_t pc = v * r; const bool c = pc <= t;
            if (c) lo = pc; else hi = pc;
            m = (a ? 4 : 0) + (bq ? 2 : 0) + (c ? 1 : 0);
        } else {
            bool open = true;
#pragma unroll
            for (int q = 0; q < K - 1; q++) {
                const uint32_t pr = lp[q] * r;
                const bool le = pr <= t;
                if (le) { lo = pr; m++; }
                else if (open) { hi = pr; open = false; }
            }
        }
        JSP_PT(2)
        const int mL = __shfl_sync(FULLMASK, m, L);
        const uint32_t lo_abs = __shfl_sync(FULLMASK, br + lo, L), width = __shfl_sync(FULLMASK, hi - lo, L);
        JSP_PT(3)
        consume(lo_abs, width);
        JSP_PT(4)
        tot += step;
        {
            const uint32_t add = lane == L ? step : 0u;        // branch-free: divergent regions cost ~50 cycles each here
#pragma unroll
            for (int q = 0; q < K; q++) lp[q] += q >= mL ? add : 0u;
            base += lane > L ? step : 0u;
        }
        bool all = false;
        if (tot > RC_BOT) {                                    // :70-77 / :113-127
            uint32_t prev = 0, s = 0;
#pragma unroll
            for (int q = 0; q < K; q++) { const uint32_t c = ((lp[q] - prev) >> 1) + 1; prev = lp[q]; s += c; lp[q] = s; }
            uint32_t incl = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += o; }
            base = incl - s;
            tot = __shfl_sync(FULLMASK, incl, 31);
            all = true;
        }
        if (all || lane == L) {
            uint4 *o4 = reinterpret_cast<uint4 *>(tab) + lane * (K / 4);
#pragma unroll
            for (int q = 0; q < K / 4; q++) o4[q] = make_uint4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]);
        }
        if (all || lane > L) tab[32 * K + lane] = base;
        // every lane writes the (identical) total: a lane only ever reads back what it wrote itself, no barrier needed
        tab[32 * K + 32] = tot;
        JSP_PT(5)
        return L * K + mL;
    }

    // copies between a cache slot and the row's home in HBM: 73 x 16 bytes, three per lane
    __device__ __forceinline__ void row_writeback(int slot, int tag)
    {
        const int lane = (int)lane_id();
        const uint4 *s4 = reinterpret_cast<const uint4 *>(&cache[slot]);
        uint4 *g4 = reinterpret_cast<uint4 *>(rows + (size_t)tag * RC_ROW_STRIDE);
#pragma unroll
        for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; if (i < (int)(sizeof(RcBig<8>) / 16)) g4[i] = s4[i]; }
    }
    __device__ __forceinline__ void row_fill(int slot, int cxi)
    {
        const int lane = (int)lane_id();
        uint4 *s4 = reinterpret_cast<uint4 *>(&cache[slot]);
        const uint32_t *grow = rows + (size_t)cxi * RC_ROW_STRIDE;
        const uint4 *g4 = reinterpret_cast<const uint4 *>(grow);
        uint4 v[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; v[k] = i < (int)(sizeof(RcBig<8>) / 16) ? g4[i] : make_uint4(0, 0, 0, 0); }
        const bool fresh = grow[32 * 8 + 33] != gen;             // not touched since the last renewI: all counts are 1
        if (fresh) {
            RcBig<8> &t = cache[slot];
#pragma unroll
            for (int q = 0; q < 8; q++) t.lp[lane * 8 + q] = q + 1;
            t.base[lane] = 8 * lane;
            if (lane == 0) { t.total = 256; t.tag = gen; t.pad[0] = t.pad[1] = 0; }
        } else {
#pragma unroll
            for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; if (i < (int)(sizeof(RcBig<8>) / 16)) s4[i] = v[k]; }
        }
        __syncwarp();
    }
    // the shared-memory slot that holds colour row cxi (loading it, and evicting the least recently used row, if needed)
    __device__ __forceinline__ uint32_t *row_slot(int cxi)
    {
        const int lane = (int)lane_id();
        const uint32_t hit = __ballot_sync(FULLMASK, lane < RC_CACHE_ROWS && my_tag == cxi);
        tick++;
        int slot;
        if (hit) {
            slot = __ffs(hit) - 1;
        } else {
            const uint32_t key = lane < RC_CACHE_ROWS ? ((my_age << 4) | (uint32_t)lane) : 0xFFFFFFFFu;
            slot = (int)(__reduce_min_sync(FULLMASK, key) & 15u);
            const int old = __shfl_sync(FULLMASK, my_tag, slot);
            if (old >= 0) row_writeback(slot, old);
            __syncwarp();
            row_fill(slot, cxi);
            if (lane == slot) my_tag = cxi;
        }
        if (lane == slot) my_age = tick;
        return cache[slot].lp;
    }
    __device__ void flush_rows()
    {
        for (int s = 0; s < RC_CACHE_ROWS; s++) {
            const int t = __shfl_sync(FULLMASK, my_tag, s);
            if (t >= 0) row_writeback(s, t);
        }
        my_tag = -1; my_age = 0;
        __syncwarp();
    }

    __device__ int decodeClr(int cxi)                                            // DecodeValUni (RangeCoder.hx:82-130) on a cached colour row
    {
        return decode_big<8>(row_slot(cxi), 400u);
    }
    __device__ int decodeN(int ptype) { return decode_big<8>(sm->ntab[ptype].lp, 400u); }   // EntroCoders.hx:142-144
    __device__ int decodeP(int ptype) { return decode_tiny<6>(sm->ptypetab[ptype], 1000u); }
    __device__ int decodeX() { return decode_big<8>(sm->xxtab.lp, 1u); }
    __device__ int decodeBT() { return decode_tiny<5>(sm->bttab, 10u); }
    __device__ int decodeBN() { return decode_big<8>(sm->ntab2.lp, 20u); }
    __device__ int decodeSXY(int n) { return decode_tiny<16>(sm->sxytab[n], 100u); }
    __device__ int decodeMX() { return decode_big<16>(sm->mvtab[0].lp, 100u); }
    __device__ int decodeMY() { return decode_big<16>(sm->mvtab[1].lp, 100u); }
    __device__ bool decodeBool() { return false; }
};

// one frame of one range-coder stream; `sm` = this warp's shared memory
__device__ __forceinline__ void sp_rc_run(const SpJob &J, RcShared &shm, uint32_t *ring, uint32_t *ptile)
{
    RcState *st = reinterpret_cast<RcState *>(J.state);
    const int lane = (int)lane_id();
    RcSmall &sm = shm.small;
    RcCoder ec;
    ec.sm = &sm; ec.cache = shm.cache; ec.my_tag = -1; ec.my_age = 0; ec.tick = 0;
    ec.rows = st->rows; ec.gen = st->gen;
    ec.fail = false; ec.poisoned = false; ec.range = 0; ec.code = 0; ec.data = J.src; ec.len = J.len; ec.pos = 0; ec.wbase = 0x80000000u; ec.nsym = 0;
#ifdef JSP_PROFILE_SECTIONS
    for (int k = 0; k < 8; k++) ec.prof[k] = 0;
#endif
    // models persist from frame to frame until the next I frame: restore the small tables
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(&st->small);
        uint4 *s = reinterpret_cast<uint4 *>(&sm);
        for (int i = lane; i < (int)(sizeof(RcSmall) / 16); i += 32) s[i] = g[i];
    }
    __syncwarp();
    uint32_t bits = 0;
    if (J.flags & SPJ_RENEW) {
        ec.renewI();
    } else if (J.flags & SPJ_IFRAME) {
        sp_decode_iframe(ec, J, ring);
        bits |= ST_CHANGED;
    } else {
        sp_decode_pframe(ec, J, bits, ptile);
    }
    ec.flush_rows();
    if (ec.failed()) {
        bits = ST_ERROR;
        if (!(J.flags & SPJ_RENEW)) sp_undo_frame(J, (J.flags & SPJ_IFRAME) != 0);
    }
    __syncwarp();
    {
        uint4 *g = reinterpret_cast<uint4 *>(&st->small);
        const uint4 *s = reinterpret_cast<const uint4 *>(&sm);
        for (int i = lane; i < (int)(sizeof(RcSmall) / 16); i += 32) g[i] = s[i];
    }
    if (lane == 0) { st->gen = ec.gen; if (bits) atomicOr(J.status, bits); if (J.symbols) *J.symbols = ec.nsym; }
    sp_signal_done(J);
#ifdef JSP_PROFILE_SECTIONS
    if (lane == 0) for (int k = 0; k < 6; k++) atomicAdd(&g_rc_prof[k], (unsigned long long)ec.prof[k]);
#endif
}

}  // namespace jsp
