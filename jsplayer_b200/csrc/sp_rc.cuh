// sp_rc.cuh -- ScreenPressor v2 entropy decode on sm_100a: 32-bit byte-wise range decoder with adaptive
// frequency tables.  Replaces reference src/RangeCoder.hx (whole file) and EntroCoderRC
// (src/EntroCoders.hx:31-180) for one stream per warp.
//
//  * coder state (range, code, read position) is replicated in every lane: no broadcasts on the serial chain;
//  * the non-colour tables (ntab 6x257, ptypetab 6x7, xxtab, ntab2, bttab, sxytab 4x17, mvtab 2x513 -- 15 KB in
//    the prefix layout below) live in shared memory for the whole frame and are saved to / restored from the
//    stream's state in HBM; the bitstream is read through a 128-byte shared-memory window;
//  * a colour context is one 1280-byte row in HBM/L2 (lane-local prefix sums of 256 counts + 32 lane bases +
//    total + generation tag; the reference's 16 group sums are derived data and are not stored).  The warp loads
//    the row with one coalesced access and finds the symbol with ONE __ballot_sync -- no prefix scan, no linear
//    search (RangeCoder.hx:58-65, :90-108) and a single division per symbol (see "table layout");
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

struct RcState {                                               // per stream, in HBM
    RcSmall small;
    uint32_t gen;                                              // generation of the colour rows (bumped by renewI)
    uint32_t pad[3];
    uint32_t *rows;                                            // RC_ROWS * RC_ROW_STRIDE u32, separately allocated
};

// floor(a / b) for 5 <= b <= 2^17 (a table total), rb = fl(1 / b).  The float estimate is within 2^10 / b + 1 of the
// quotient, so |rem| < 2^24 is exact in float and one refinement lands within one of the quotient; the two
// predicated corrections make it exact (the trailing loops never run; they keep the result exact by construction).
__device__ __forceinline__ uint32_t udiv_small(uint32_t a, uint32_t b, float rb)
{
    uint32_t q = __float2uint_rz(__uint2float_rz(a) * rb);
    int32_t rem = (int32_t)(a - q * b);
    const int32_t adj = __float2int_rd(__int2float_rn(rem) * rb);
    q += (uint32_t)adj; rem -= adj * (int32_t)b;
    if (rem < 0) { q--; rem += (int32_t)b; }
    if (rem >= (int32_t)b) { q++; rem -= (int32_t)b; }
    while (rem < 0) { q--; rem += (int32_t)b; }
    while (rem >= (int32_t)b) { q++; rem -= (int32_t)b; }
    return q;
}

struct RcCoder {
    static constexpr bool kCanDecodeBool = false;              // EntroCoders.hx:178
    RcSmall *sm;                                               // shared memory
    uint32_t *rows;
    uint32_t gen;
    uint32_t range, code;
    const uint8_t *data;
    uint32_t len, pos, wbase;
    uint32_t nsym;                                             // symbols decoded in this frame (reporting only)
    bool poisoned, fail;

    __device__ __forceinline__ bool failed() const { return fail; }

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
        if (poisoned) fail = true;
        const uint32_t r = udiv_small(range, tot, __frcp_rn(__uint2float_rn(tot)));
        const uint32_t codev = poisoned ? 0u : code;
        const uint32_t pr = p * r;
        const int s = __popc(__ballot_sync(FULLMASK, lane < N && pr <= codev));
        if (s >= N) { range = r; fail = true; return N - 1; }
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

    // RangeCoder.hx:51-80 (256 / 512 symbols) and :82-130 (colour rows; the 16 group sums are derived data)
    template <int K, bool IS_ROW>
    __device__ __forceinline__ int decode_big(uint32_t *tab, uint32_t step)
    {
        const int lane = (int)lane_id();
        nsym++;
        uint32_t lp[K];
        uint32_t base, tot;
        bool fresh = false;
        {
            const uint4 *t4 = reinterpret_cast<const uint4 *>(tab) + lane * (K / 4);
#pragma unroll
            for (int q = 0; q < K / 4; q++) { const uint4 v = t4[q]; lp[4 * q] = v.x; lp[4 * q + 1] = v.y; lp[4 * q + 2] = v.z; lp[4 * q + 3] = v.w; }
            base = tab[32 * K + lane];
            if (IS_ROW) {
                const uint2 meta = *reinterpret_cast<const uint2 *>(tab + 32 * K + 32);   // total, generation tag
                tot = meta.x;
                fresh = meta.y != gen;                         // not touched since the last renewI: all counts are 1
                if (fresh) {
#pragma unroll
                    for (int q = 0; q < K; q++) lp[q] = q + 1;
                    base = K * lane; tot = 32 * K;
                }
            } else tot = tab[32 * K + 32];
        }
        if (poisoned) fail = true;
        const uint32_t r = udiv_small(range, tot, __frcp_rn(__uint2float_rn(tot)));
        const uint32_t codev = poisoned ? 0u : code;
        const uint32_t br = base * r;
        if (codev >= tot * r) { range = r; fail = true; return 32 * K - 1; }     // value >= total: not a valid stream
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
        const int mL = __shfl_sync(FULLMASK, m, L);
        const uint32_t lo_abs = __shfl_sync(FULLMASK, br + lo, L), width = __shfl_sync(FULLMASK, hi - lo, L);
        consume(lo_abs, width);
        tot += step;
        if (lane == L) {
#pragma unroll
            for (int q = 0; q < K; q++) if (q >= mL) lp[q] += step;
        }
        if (lane > L) base += step;
        bool all = fresh;
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
        if (IS_ROW) *reinterpret_cast<uint2 *>(tab + 32 * K + 32) = make_uint2(tot, gen);
        else tab[32 * K + 32] = tot;
        return L * K + mL;
    }

    __device__ int decodeClr(int cxi)                                            // DecodeValUni on a colour row in HBM/L2
    {
        return decode_big<8, true>(rows + (size_t)cxi * RC_ROW_STRIDE, 400u);
    }
    __device__ int decodeN(int ptype) { return decode_big<8, false>(sm->ntab[ptype].lp, 400u); }   // EntroCoders.hx:142-144
    __device__ int decodeP(int ptype) { return decode_tiny<6>(sm->ptypetab[ptype], 1000u); }
    __device__ int decodeX() { return decode_big<8, false>(sm->xxtab.lp, 1u); }
    __device__ int decodeBT() { return decode_tiny<5>(sm->bttab, 10u); }
    __device__ int decodeBN() { return decode_big<8, false>(sm->ntab2.lp, 20u); }
    __device__ int decodeSXY(int n) { return decode_tiny<16>(sm->sxytab[n], 100u); }
    __device__ int decodeMX() { return decode_big<16, false>(sm->mvtab[0].lp, 100u); }
    __device__ int decodeMY() { return decode_big<16, false>(sm->mvtab[1].lp, 100u); }
    __device__ bool decodeBool() { return false; }
};

// one frame of one range-coder stream; `sm` = this warp's shared memory
__device__ __forceinline__ void sp_rc_run(const SpJob &J, RcSmall &sm)
{
    RcState *st = reinterpret_cast<RcState *>(J.state);
    const int lane = (int)lane_id();
    RcCoder ec;
    ec.sm = &sm; ec.rows = st->rows; ec.gen = st->gen;
    ec.fail = false; ec.poisoned = false; ec.range = 0; ec.code = 0; ec.data = J.src; ec.len = J.len; ec.pos = 0; ec.wbase = 0x80000000u; ec.nsym = 0;
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
        sp_decode_iframe(ec, J);
        bits |= ST_CHANGED;
    } else {
        sp_decode_pframe(ec, J, bits);
    }
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
}

}  // namespace jsp
