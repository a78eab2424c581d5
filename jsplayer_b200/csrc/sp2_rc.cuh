// sp2_rc.cuh -- ScreenPressor v2 entropy decode on sm_100a, second generation: the serial chain of one symbol cut to
// the minimum.  Replaces reference src/RangeCoder.hx (whole file) and EntroCoderRC (src/EntroCoders.hx:31-180) for one
// stream per warp, bit-exact with oracle/rangecoder_oracle.c (including its defined failure behaviour).
//
// Round 1's decoder (sp_rc.cuh, kept for A/B runs: JSP_SP_GEN=1) spent ~740 cycles per symbol; its section profile
// (profiles/r01_sp_section_profile_final.txt) put 160 of them into a floating-point reciprocal, ~110 into ballot +
// search + three shuffles, ~110 into the update and ~450 into a 6-symbol table's decode.  What changed:
//
//  * NO division on the chain.  Every table carries the fixed-point reciprocal of its total, inv = floor((2^32-1)/tot).
//    It is computed when the total CHANGES -- i.e. during the previous decode of that table, from tot + step, which is
//    known before the search starts, so the ~25-instruction integer division overlaps with the symbol search instead of
//    preceding it.  r = range / tot is then umulhi + multiply + one compare: with that inv the estimate is never above
//    and at most one below the true quotient (a < 2^32, proof at udiv1()).
//  * Tables of <= 16 symbols (ptypetab, bttab, sxytab) are decoded by EVERY lane from broadcast shared-memory loads:
//    no ballot, no shuffle, no divergence; every lane computes and stores identical values, so no barrier either.
//  * 256 / 512-symbol tables keep round 1's layout (per lane K inclusive prefix sums + the lane's exclusive base), but a
//    lane loads only its base: one ballot finds the owning lane L, then EVERY lane loads L's K prefixes (one broadcast
//    address) and runs the same 3-probe search -- the symbol, its interval and the update are warp-uniform values, the
//    three shuffles are gone and the per-symbol shared-memory traffic drops from 1.3 KB to ~200 B.
//  * The bitstream is read through a register: `buf` holds the aligned 32-bit word of the next byte, the following word
//    is requested from the shared-memory window as soon as a word is finished, and the window itself (128 B) is
//    refilled from a register prefetched one window ahead -- no load of any kind sits between a symbol and its bytes.
#pragma once
#include "sp_common.cuh"
#include <cstddef>
#include <cstring>

namespace jsp {
namespace g2 {

constexpr uint32_t RC_TOP = 0x01000000u, RC_BOT = 0x010000u;   // RangeCoder.hx:12-13

// floor(a / b) for any a < 2^32 and 1 <= b < 2^31 given inv = floor((2^32 - 1) / b):
// inv = (2^32 - 1 - e) / b with 0 <= e < b, so a * inv / 2^32 = a / b - a (1 + e) / (b 2^32) and the subtrahend is below
// (1 + e) / b <= 1: the estimate q = umulhi(a, inv) is floor(a / b) or one less, never more.
__device__ __forceinline__ uint32_t udiv1(uint32_t a, uint32_t b, uint32_t inv)
{
    const uint32_t q = __umulhi(a, inv);
    const uint32_t rem = a - q * b;
    return q + (rem >= b ? 1u : 0u);
}
__device__ __forceinline__ uint32_t recip32(uint32_t b) { return 0xFFFFFFFFu / b; }

// ---- table layouts --------------------------------------------------------------------------------------------
template <int K>
struct RcBig {                                                 // K = 8: 256 symbols, K = 16: 512 symbols
    uint32_t lp[32 * K];                                       // lane-local inclusive prefix sums of the lane's K counts
    uint32_t base[32];                                         // sum of all counts of lower lanes
    uint32_t total;
    uint32_t tag;                                              // colour rows: generation; shared-memory tables: unused
    uint32_t inv;                                              // floor((2^32 - 1) / total)
    uint32_t pad;
};
template <int NW>
struct alignas(16) RcTiny {                                    // N symbols: P[0..N-1] inclusive cumulative counts, P[NW-1] = inv
    uint32_t P[NW];
};
constexpr int RC_ROW_STRIDE = 320;                             // u32 per colour row in HBM (RcBig<8> = 292, padded to 1280 B)
constexpr int RC_ROWS = 3 * 4096;
static_assert(sizeof(RcBig<8>) == 292 * 4, "row layout");

struct RcSmall {
    RcBig<8> ntab[6], xxtab, ntab2;
    RcBig<16> mvtab[2];
    RcTiny<20> sxytab[4];
    RcTiny<8> ptypetab[6], bttab;
};

constexpr int RC_CACHE_ROWS = 12;                              // LRU cache of colour rows in shared memory (see sp_rc.cuh)
struct RcShared {
    RcSmall small;
    RcBig<8> cache[RC_CACHE_ROWS];
    alignas(16) uint8_t win[128];                              // bitstream window
};

struct RcState {                                               // per stream, in HBM
    RcSmall small;
    uint32_t gen;                                              // generation of the colour rows (bumped by renewI)
    uint32_t pad[3];
    uint32_t *rows;                                            // RC_ROWS * RC_ROW_STRIDE u32, separately allocated
};

// cold: apply the pending +step, then every count -> (count >> 1) + 1 (RangeCoder.hx:70-77 / :113-127), prefixes and bases
// rebuilt.  A free function on purpose: a non-inlined MEMBER would force the whole coder object into local memory.
template <int K>
static __device__ __noinline__ void rc_rescale_big(uint32_t *tab, int L, int m, uint32_t step)
{
    const int lane = (int)lane_id();
    __syncwarp();
    uint32_t lp[K];
#pragma unroll
    for (int q = 0; q < K; q++) lp[q] = tab[lane * K + q] + ((lane == L && q >= m) ? step : 0u);
    uint32_t prev = 0, s = 0;
#pragma unroll
    for (int q = 0; q < K; q++) { const uint32_t c = ((lp[q] - prev) >> 1) + 1; prev = lp[q]; s += c; lp[q] = s; }
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += o; }
    const uint32_t tot = __shfl_sync(FULLMASK, incl, 31);
#pragma unroll
    for (int q = 0; q < K; q++) tab[lane * K + q] = lp[q];
    tab[32 * K + lane] = incl - s;
    if (lane == 0) { tab[32 * K + 32] = tot; tab[32 * K + 34] = 0xFFFFFFFFu / tot; }
    __syncwarp();
}

static __device__ __forceinline__ void rc_row_writeback(const RcBig<8> *cache, uint32_t *rows, int slot, int tag)
{
    const int lane = (int)lane_id();
    const uint4 *s4 = reinterpret_cast<const uint4 *>(&cache[slot]);
    uint4 *g4 = reinterpret_cast<uint4 *>(rows + (size_t)tag * RC_ROW_STRIDE);
#pragma unroll
    for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; if (i < (int)(sizeof(RcBig<8>) / 16)) g4[i] = s4[i]; }
}
// cold: evict the least recently used row, load row cxi (or build it: rows not touched since the last renewI are all ones)
static __device__ __noinline__ int rc_row_miss(RcBig<8> *cache, uint32_t *rows, uint32_t gen, int my_tag, uint32_t my_age, int cxi)
{
    const int lane = (int)lane_id();
    const uint32_t key = lane < RC_CACHE_ROWS ? ((my_age << 4) | (uint32_t)lane) : 0xFFFFFFFFu;
    const int slot = (int)(__reduce_min_sync(FULLMASK, key) & 15u);
    const int old = __shfl_sync(FULLMASK, my_tag, slot);
    if (old >= 0) rc_row_writeback(cache, rows, slot, old);
    __syncwarp();
    uint4 *s4 = reinterpret_cast<uint4 *>(&cache[slot]);
    const uint32_t *grow = rows + (size_t)cxi * RC_ROW_STRIDE;
    const uint4 *g4 = reinterpret_cast<const uint4 *>(grow);
    uint4 v[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; v[k] = i < (int)(sizeof(RcBig<8>) / 16) ? g4[i] : make_uint4(0, 0, 0, 0); }
    const bool fresh = grow[32 * 8 + 33] != gen;
    if (fresh) {
        RcBig<8> &t = cache[slot];
#pragma unroll
        for (int q = 0; q < 8; q++) t.lp[lane * 8 + q] = q + 1;
        t.base[lane] = 8 * lane;
        if (lane == 0) { t.total = 256; t.tag = gen; t.inv = 0xFFFFFFFFu / 256u; t.pad = 0; }
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; if (i < (int)(sizeof(RcBig<8>) / 16)) s4[i] = v[k]; }
    }
    __syncwarp();
    return slot;
}

struct RcCoder {
    static constexpr bool kCanDecodeBool = false;              // EntroCoders.hx:178
    RcSmall *sm;
    RcBig<8> *cache;
    uint8_t *win;
    int my_tag;                                                // lane < RC_CACHE_ROWS: context index held by slot `lane`, -1 = empty
    uint32_t my_age, tick;
    uint32_t *rows;
    uint32_t gen;
    uint32_t range, code;
    const uint8_t *data;
    uint32_t len, pos, wbase;                                  // pos = index of the next byte; wbase = first byte of the window
    uint32_t buf;                                              // the aligned word that holds byte `pos`
    uint32_t pre;                                              // this lane's word of the NEXT window (prefetched)
    uint32_t nsym;
    bool poisoned, fail;

    __device__ __forceinline__ bool failed() const { return fail; }
    __device__ __forceinline__ void fail_frame() { fail = true; range = 0; }   // see decode_tiny: a zero range fails every later symbol

    // ---- bitstream ----
    __device__ __forceinline__ uint32_t load_window_word(uint32_t wb) const    // this lane's 4 bytes of the window at wb
    {
        const uint32_t p = wb + 4u * lane_id();
        uint32_t w = 0;
        if (p + 4u <= len && ((reinterpret_cast<uintptr_t>(data) + p) & 3u) == 0) return __ldg(reinterpret_cast<const uint32_t *>(data + p));
#pragma unroll
        for (int k = 0; k < 4; k++) if (p + k < len) w |= (uint32_t)__ldg(data + p + k) << (8 * k);
        return w;
    }
    __device__ __forceinline__ void fetch_word()                               // buf = word holding byte `pos` (pos is 4-aligned here)
    {
        if (pos - wbase >= 128u) {                                             // next window: comes out of the prefetch register
            __syncwarp();
            wbase += 128u;
            reinterpret_cast<uint32_t *>(win)[lane_id()] = pre;
            pre = load_window_word(wbase + 128u);
            __syncwarp();
        }
        buf = reinterpret_cast<const uint32_t *>(win)[(pos - wbase) >> 2];
    }
    __device__ __forceinline__ void next_byte()                                // RangeCoder.hx:41 `code = code * 256 + data[pos++]`
    {
        if (pos >= len) poisoned = true;                                       // JS: code becomes NaN for good
        code = (code << 8) | ((buf >> ((pos & 3u) * 8u)) & 0xFFu);
        pos++;
        if ((pos & 3u) == 0) fetch_word();
    }
    __device__ __forceinline__ void decodeBegin(const uint8_t *src, uint32_t n, uint32_t pos0)   // RangeCoder.hx:19-34
    {
        data = src; len = n; code = 0; range = 0xFFFFFFFFu; poisoned = false;
        pos = pos0 + 1;
        __syncwarp();
        wbase = pos & ~127u;
        reinterpret_cast<uint32_t *>(win)[lane_id()] = load_window_word(wbase);
        pre = load_window_word(wbase + 128u);
        __syncwarp();
        buf = reinterpret_cast<const uint32_t *>(win)[(pos - wbase) >> 2];
        next_byte(); next_byte(); next_byte(); next_byte();
    }
    // RangeCoder.hx:36-43 with the products already formed: lo = cumFreq * r, width = freq * r
    __device__ __forceinline__ void consume(uint32_t lo, uint32_t width)
    {
        code -= lo;
        range = width;
        while (range < RC_TOP) { next_byte(); range <<= 8; }
    }

    // ---- model reset ----
    template <int K>
    __device__ __forceinline__ void init_big(RcBig<K> &t)
    {
        const int lane = (int)lane_id();
#pragma unroll
        for (int q = 0; q < K; q++) t.lp[lane * K + q] = q + 1;
        t.base[lane] = K * lane;
        if (lane == 0) { t.total = 32 * K; t.tag = 0; t.inv = recip32(32 * K); t.pad = 0; }
    }
    template <int N, int NW>
    __device__ __forceinline__ void init_tiny(RcTiny<NW> &t)
    {
        const int lane = (int)lane_id();
        if (lane < NW) t.P[lane] = lane < N ? (uint32_t)lane + 1u : (lane == NW - 1 ? recip32(N) : 0u);
    }
    __device__ __forceinline__ void renewI()                                                     // EntroCoders.hx:81-130
    {
        gen = gen + 1;
        for (int t = 0; t < 6; t++) { init_big(sm->ntab[t]); init_tiny<6>(sm->ptypetab[t]); }
        init_big(sm->xxtab); init_big(sm->ntab2);
        init_big(sm->mvtab[0]); init_big(sm->mvtab[1]);
        for (int t = 0; t < 4; t++) init_tiny<16>(sm->sxytab[t]);
        init_tiny<5>(sm->bttab);
        __syncwarp();
    }

    // RangeCoder.hx:51-80 for tables of N <= 16 symbols: every lane runs the whole decode on broadcast loads.
    template <int N, int NW>
    __device__ __forceinline__ int decode_tiny(RcTiny<NW> &t, uint32_t step)
    {
        nsym++;
        uint32_t P[NW];
        {
            const uint4 *t4 = reinterpret_cast<const uint4 *>(t.P);
#pragma unroll
            for (int q = 0; q < NW / 4; q++) { const uint4 v = t4[q]; P[4 * q] = v.x; P[4 * q + 1] = v.y; P[4 * q + 2] = v.z; P[4 * q + 3] = v.w; }
        }
        const uint32_t tot = P[N - 1];
        const uint32_t inv_next = recip32(tot + step);                          // off the chain: overlaps with the search
        // A failed frame decodes nothing more (the models stay as they were), at no cost to the symbol chain: failing
        // zeroes `range`, a zero range gives r = 0, every product is then 0 <= code and the search runs off the table
        // again.  A symbol asked for after the data ran out (`poisoned`) fails the same way.
        if (poisoned) range = 0;
        const uint32_t r = udiv1(range, tot, P[NW - 1]);
        const uint32_t codev = code;
        // products ascend: lo = the largest one <= code, hi = the smallest one above it (balanced max / min trees, not a chain)
        bool le[N];
        uint32_t vlo[16], vhi[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { vlo[i] = 0u; vhi[i] = 0xFFFFFFFFu; }
        int s = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t pr = P[i] * r;
            le[i] = pr <= codev;
            vlo[i] = le[i] ? pr : 0u;
            vhi[i] = le[i] ? 0xFFFFFFFFu : pr;
            s += le[i] ? 1 : 0;
        }
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
            for (int i = 0; i < w; i++) { vlo[i] = max(vlo[i], vlo[i + w]); vhi[i] = min(vhi[i], vhi[i + w]); }
        const uint32_t lo = vlo[0], hi = vhi[0];
        if (s >= N) { range = 0; fail = true; return N - 1; }
        consume(lo, hi - lo);
#pragma unroll
        for (int i = 0; i < N; i++) P[i] += le[i] ? 0u : step;                  // cumulative counts of symbols >= s
        uint32_t ninv = inv_next;
        if (tot + step > RC_BOT) {                                               // :70-77: every count -> (count >> 1) + 1
            uint32_t prev = 0, acc = 0;
#pragma unroll
            for (int i = 0; i < N; i++) { const uint32_t c = ((P[i] - prev) >> 1) + 1; prev = P[i]; acc += c; P[i] = acc; }
            ninv = recip32(acc);
        }
        P[NW - 1] = ninv;
        {
            uint4 *t4 = reinterpret_cast<uint4 *>(t.P);                          // every lane stores the same values: no barrier
#pragma unroll
            for (int q = 0; q < NW / 4; q++) t4[q] = make_uint4(P[4 * q], P[4 * q + 1], P[4 * q + 2], P[4 * q + 3]);
        }
        return s;
    }

    // RangeCoder.hx:51-80 (256 / 512 symbols) and :82-130 (colour rows; the reference's 16 group sums are derived data and
    // are not kept) on a table in shared memory.
    template <int K>
    __device__ __forceinline__ int decode_big(uint32_t *tab, uint32_t step)
    {
        const int lane = (int)lane_id();
        nsym++;
        const uint32_t base = tab[32 * K + lane];
        const uint4 hdr = *reinterpret_cast<const uint4 *>(tab + 32 * K + 32);   // total, tag, inv, pad
        const uint32_t tot = hdr.x;
        const uint32_t inv_next = recip32(tot + step);                          // off the chain
        if (poisoned) range = 0;                                                 // as in decode_tiny: fails below, touches nothing
        const uint32_t r = udiv1(range, tot, hdr.z);
        const uint32_t codev = code;
        const uint32_t br = base * r;
        if (codev >= tot * r) { range = 0; fail = true; return 32 * K - 1; }    // value >= total: not a valid stream
        const int L = __popc(__ballot_sync(FULLMASK, br <= codev)) - 1;          // lane 0 has base 0: L >= 0
        // every lane now works on lane L's K prefixes (one broadcast address) and on L's base product
        uint32_t lp[K];
        {
            const uint4 *t4 = reinterpret_cast<const uint4 *>(tab) + L * (K / 4);
#pragma unroll
            for (int q = 0; q < K / 4; q++) { const uint4 v = t4[q]; lp[4 * q] = v.x; lp[4 * q + 1] = v.y; lp[4 * q + 2] = v.z; lp[4 * q + 3] = v.w; }
        }
        const uint32_t brL = __shfl_sync(FULLMASK, br, L);
        const uint32_t t = codev - brL;
        // the symbol's inclusive prefix exceeds t (the next lane's base is above the value), so m = #{q : lp[q] * r <= t} is
        // at most K - 1: a binary search over lp[0 .. K-2] finds it together with lo = lp[m-1] * r (0 if m = 0), hi = lp[m] * r
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
            // K == 16: 4 probes
            const uint32_t pa = lp[7] * r; const bool a = pa <= t;
            if (a) lo = pa; else hi = pa;
            const uint32_t pb = (a ? lp[11] : lp[3]) * r; const bool bq = pb <= t;
            if (bq) lo = pb; else hi = pb;
            const uint32_t vc = bq ? (a ? lp[13] : lp[5]) : (a ? lp[9] : lp[1]);
            const uint32_t pc = vc * r; const bool c = pc <= t;
            if (c) lo = pc; else hi = pc;
            const int i3 = (a ? 8 : 0) + (bq ? 4 : 0) + (c ? 2 : 0);             // probe lp[i3]
            uint32_t vd = lp[0];
#pragma unroll
            for (int q = 0; q < 16; q += 2) vd = (i3 == q) ? lp[q] : vd;
            const uint32_t pd = vd * r; const bool d = pd <= t;
            if (d) lo = pd; else hi = pd;
            m = i3 + (d ? 1 : 0);
        }
        consume(brL + lo, hi - lo);
        uint32_t ntot = tot + step;
        if (ntot > RC_BOT) {                                                     // :70-77 / :113-127: rescale the whole table
            rc_rescale_big<K>(tab, L, m, step);
            return L * K + m;
        }
        {
#pragma unroll
            for (int q = 0; q < K; q++) lp[q] += q >= m ? step : 0u;
            uint4 *o4 = reinterpret_cast<uint4 *>(tab) + L * (K / 4);            // every lane stores the same values
#pragma unroll
            for (int q = 0; q < K / 4; q++) o4[q] = make_uint4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]);
            if (lane > L) tab[32 * K + lane] = base + step;                      // a lane reads back only its own base
            *reinterpret_cast<uint4 *>(tab + 32 * K + 32) = make_uint4(ntot, hdr.y, inv_next, hdr.w);
        }
        return L * K + m;
    }
    // ---- colour-row cache (as round 1: 12 fully associative LRU slots, rows decoded in shared memory) ----
    __device__ __forceinline__ uint32_t *row_slot(int cxi)
    {
        const int lane = (int)lane_id();
        const uint32_t hit = __ballot_sync(FULLMASK, lane < RC_CACHE_ROWS && my_tag == cxi);
        tick++;
        int slot;
        if (hit) slot = __ffs(hit) - 1;
        else {
            slot = rc_row_miss(cache, rows, gen, my_tag, my_age, cxi);
            if (lane == slot) my_tag = cxi;
        }
        if (lane == slot) my_age = tick;
        return cache[slot].lp;
    }
    __device__ __forceinline__ void flush_rows()
    {
        for (int s = 0; s < RC_CACHE_ROWS; s++) {
            const int t = __shfl_sync(FULLMASK, my_tag, s);
            if (t >= 0) rc_row_writeback(cache, rows, s, t);
        }
        my_tag = -1; my_age = 0;
        __syncwarp();
    }

    __device__ __forceinline__ int decodeClr(int cxi) { return decode_big<8>(row_slot(cxi), 400u); }    // DecodeValUni, RangeCoder.hx:82-130
    __device__ __forceinline__ int decodeN(int ptype) { return decode_big<8>(sm->ntab[ptype].lp, 400u); }   // EntroCoders.hx:142-144
    __device__ __forceinline__ int decodeP(int ptype) { return decode_tiny<6>(sm->ptypetab[ptype], 1000u); }
    __device__ __forceinline__ int decodeX() { return decode_big<8>(sm->xxtab.lp, 1u); }
    __device__ __forceinline__ int decodeBT() { return decode_tiny<5>(sm->bttab, 10u); }
    __device__ __forceinline__ int decodeBN() { return decode_big<8>(sm->ntab2.lp, 20u); }
    __device__ __forceinline__ int decodeSXY(int n) { return decode_tiny<16>(sm->sxytab[n], 100u); }
    __device__ __forceinline__ int decodeMX() { return decode_big<16>(sm->mvtab[0].lp, 100u); }
    __device__ __forceinline__ int decodeMY() { return decode_big<16>(sm->mvtab[1].lp, 100u); }
    __device__ __forceinline__ bool decodeBool() { return false; }

    // ---- per-frame set-up / tear-down: the small tables travel between the stream's state in HBM and shared memory ----
    __device__ __forceinline__ void open(const SpJob &J, RcShared &shm)
    {
        RcState *st = reinterpret_cast<RcState *>(J.state);
        sm = &shm.small; cache = shm.cache; win = shm.win;
        my_tag = -1; my_age = 0; tick = 0;
        rows = st->rows; gen = st->gen;
        fail = false; poisoned = false; range = 0; code = 0; data = J.src; len = J.len; pos = 0; wbase = 0; buf = 0; pre = 0; nsym = 0;
        const uint4 *g = reinterpret_cast<const uint4 *>(&st->small);
        uint4 *s = reinterpret_cast<uint4 *>(&shm.small);
        for (int i = (int)lane_id(); i < (int)(sizeof(RcSmall) / 16); i += 32) s[i] = g[i];
        __syncwarp();
    }
    __device__ __forceinline__ void close(const SpJob &J, RcShared &shm)
    {
        RcState *st = reinterpret_cast<RcState *>(J.state);
        flush_rows();
        uint4 *g = reinterpret_cast<uint4 *>(&st->small);
        const uint4 *s = reinterpret_cast<const uint4 *>(&shm.small);
        for (int i = (int)lane_id(); i < (int)(sizeof(RcSmall) / 16); i += 32) g[i] = s[i];
        if (lane_id() == 0) st->gen = gen;
    }
};

}  // namespace g2
}  // namespace jsp
