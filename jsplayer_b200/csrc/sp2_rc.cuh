// sp2_rc.cuh -- ScreenPressor v2 entropy decode on sm_100a, second generation.  Replaces reference src/RangeCoder.hx
// (whole file) and EntroCoderRC (src/EntroCoders.hx:31-180) for one stream per warp, bit-exact with the CPU checker's
// restatement (including its defined failure behaviour).
//
// What limits this decoder is not memory and not arithmetic throughput: ONE warp decodes a stream (every symbol depends
// on the previous one through the coder state and the adaptive model), and a lone warp issues an instruction every 2+
// cycles whatever its dependencies (each pipe takes a warp instruction per 2 cycles) and pays ~20 cycles per
// data-dependent branch (predicate latency to the branch unit + the refetch).  tools/sass_timeline.py replays the SASS of
// a region with the control bits the assembler encoded and reproduces the clock64 section timings within 10 %.  So the
// cost of a symbol is its INSTRUCTION COUNT and its BRANCH COUNT; round 1 (sp_rc.cuh, kept for A/B runs: JSP_SP_GEN=1)
// spent 165 instructions and ~740 cycles per symbol.  Design rules here:
//
//  * One value per lane, warp collectives instead of per-lane loops: a 6-symbol table is 1 multiply + 1 compare + 1
//    ballot + 2 shuffles; a 256-symbol table is that twice (32 lane bases, then the 8 prefixes of the owning lane).
//  * NO division.  r = range / total needs the exact quotient; every table carries inv = floor((2^32-1) / total), which
//    makes it umulhi + multiply + one compare (proof at udiv1()).  A table's total moves by a fixed step per symbol, so
//    the reciprocals of its NEXT 32 totals are computed at once -- lane j takes total + j * step, one 17-instruction
//    division sequence for all 32 -- and the per-symbol update just picks the next one: 2 instructions instead of 17.
//  * ONE branch per symbol.  Everything rare (search ran off the table, rescale, reciprocal batch used up, bitstream
//    window half used up, end of data) is OR-ed into a single predicate after the search; the common path is straight-line.
//  * The byte-wise renormalisation loop is one step: clz(range) / 8 bytes are funnel-shifted into `code` at once out of
//    two registers holding the current and the next bitstream word; both are re-read from a 256-byte shared-memory window
//    after every symbol (no "crossed a word?" branch), and the window is refilled from a register prefetched one half ahead.
#pragma once
#include "sp_common.cuh"
#include <cstddef>
#include <cstring>

namespace jsp {
namespace g2 {

constexpr uint32_t RC_TOP = 0x01000000u, RC_BOT = 0x010000u;   // RangeCoder.hx:12-13

// Shared memory is addressed by 32-bit shared-window addresses kept in registers: every conversion of a generic pointer
// costs an S2UR + ULEA pair (~20 cycles of latency), and plain C++ loads get sunk below the first branch that does not need
// them.  The asm statements keep program order.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32_if(bool p, uint32_t a, uint32_t v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.shared.u32 [%1], %2; }" ::"r"((uint32_t)p), "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

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

// ---- table layouts (u32 words) ---------------------------------------------------------------------------------
// hdr = {total, tag, inv, cnt}: inv = recip32(total); fut[j] = recip32(total + (j - cnt) * step), i.e. the reciprocals of
// the totals this table will have after the next symbols (valid for j > cnt; refilled when cnt reaches 32 or the table is
// rescaled).  Only the first RC_ROW_WORDS words of a colour row travel to HBM; fut[] is rebuilt when a row is loaded.
template <int K>
struct alignas(16) RcBig {                                     // K = 8: 256 symbols, K = 16: 512 symbols
    uint32_t lp[32 * K];                                       // lane-local inclusive prefix sums of the lane's K counts
    uint32_t base[32];                                         // sum of all counts of lower lanes
    uint32_t hdr[4];                                           // total, tag (colour rows: generation), inv, cnt
    uint32_t fut[32];
};
struct alignas(16) RcTiny {                                    // N <= 16 symbols
    uint32_t P[32];                                            // P[lane] = inclusive cumulative count (lane < N); P[28..31] = -, cnt, total, inv
    uint32_t fut[32];
};
constexpr int RC_ROW_WORDS = 32 * 8 + 32 + 4;                  // 292: what a colour row keeps in HBM
constexpr int RC_ROW_STRIDE = 320;                             // u32 per colour row in HBM (padded to 1280 B)
constexpr int RC_ROWS = 3 * 4096;
constexpr uint32_t RC_CLR_STEP = 400;                          // SC_STEP, EntroCoders.hx:43

struct RcSmall {
    RcBig<8> ntab[6];
    RcTiny ptypetab[6];
    // tables only P frames use: an I-frame kernel keeps them out of shared memory (it resets them directly in HBM)
    RcBig<8> xxtab, ntab2;
    RcBig<16> mvtab[2];
    RcTiny sxytab[4], bttab;
};
constexpr uint32_t RC_SMALL_I_BYTES = (uint32_t)offsetof(RcSmall, xxtab);   // the part every frame type needs

constexpr int RC_CACHE_ROWS = 10;                              // LRU cache of colour rows in shared memory
struct RcShared {                                              // bitstream window and row cache first: their offsets do not depend
    alignas(16) uint8_t win[256];                              // on how much of `small` a kernel keeps (two 128-byte halves, circular)
    RcBig<8> cache[RC_CACHE_ROWS];
    RcSmall small;
};
constexpr uint32_t RC_SHARED_I_BYTES = (uint32_t)offsetof(RcShared, small) + RC_SMALL_I_BYTES;

struct RcState {                                               // per stream, in HBM
    RcSmall small;
    uint32_t gen;                                              // generation of the colour rows (bumped by renewI)
    uint32_t pad[3];
    uint32_t *rows;                                            // RC_ROWS * RC_ROW_STRIDE u32, separately allocated
};

// ---- cold paths: free functions on purpose (a non-inlined MEMBER would force the whole coder object into local memory) ----

// reciprocals of the next 32 totals, and the header for the new batch (every lane stores the same header)
static __device__ __noinline__ void rc_refill_fut(uint32_t *hdr, uint32_t *fut, uint32_t total, uint32_t tag, uint32_t step)
{
    const uint32_t v = recip32(total + lane_id() * step);
    fut[lane_id()] = v;
    const uint32_t inv = __shfl_sync(FULLMASK, v, 0);
    hdr[0] = total; hdr[1] = tag; hdr[2] = inv; hdr[3] = 0;
    __syncwarp();
}
// a big table whose total passed BOT: apply the pending +step, then every count -> (count >> 1) + 1 (RangeCoder.hx:70-77 /
// :113-127), prefixes and bases rebuilt
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
    __syncwarp();
    rc_refill_fut(tab + 32 * K + 32, tab + 32 * K + 36, tot, tab[32 * K + 33], step);
}
// a tiny table whose total passed BOT; p = this lane's cumulative count with the pending +step already applied
template <int N>
static __device__ __noinline__ void rc_rescale_tiny(uint32_t *P, uint32_t p, uint32_t step)
{
    const int lane = (int)lane_id();
    const uint32_t prev = __shfl_up_sync(FULLMASK, p, 1);
    uint32_t c = lane < N ? ((p - (lane ? prev : 0u)) >> 1) + 1 : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULLMASK, c, d); if (lane >= d) c += o; }
    const uint32_t tot = __shfl_sync(FULLMASK, c, N - 1);
    __syncwarp();
    if (lane < N) P[lane] = c;
    const uint32_t v = recip32(tot + lane * step);
    P[32 + lane] = v;
    const uint32_t inv = __shfl_sync(FULLMASK, v, 0);
    __syncwarp();
    if (lane == 0) { P[28] = 0; P[29] = 0; P[30] = tot; P[31] = inv; }
    __syncwarp();
}
static __device__ __noinline__ void rc_refill_tiny(uint32_t *P, uint32_t total, uint32_t step)
{
    const uint32_t v = recip32(total + lane_id() * step);
    P[32 + lane_id()] = v;
    const uint32_t inv = __shfl_sync(FULLMASK, v, 0);
    __syncwarp();
    if (lane_id() == 0) { P[28] = 0; P[29] = 0; P[30] = total; P[31] = inv; }
    __syncwarp();
}

// this lane's 4 bytes of the 128-byte piece of the frame at half_base (bytes past the end read as 0)
static __device__ __forceinline__ uint32_t rc_load_half_word(const uint8_t *data, uint32_t len, uint32_t half_base)
{
    const uint32_t p = half_base + 4u * lane_id();
    uint32_t w = 0;
    if (p + 4u <= len && ((reinterpret_cast<uintptr_t>(data) + p) & 3u) == 0) return __ldg(reinterpret_cast<const uint32_t *>(data + p));
#pragma unroll
    for (int k = 0; k < 4; k++) if (p + k < len) w |= (uint32_t)__ldg(data + p + k) << (8 * k);
    return w;
}
// once per 128 bytes, after `pos` has entered a new half of the window: the half it left takes the prefetched data (256 bytes
// further on), and the prefetch register is reloaded for the half after that.  Returns the new prefetch word.
static __device__ __noinline__ uint32_t rc_window_refill(uint32_t win_a, const uint8_t *data, uint32_t len, uint32_t pos, uint32_t pre)
{
    __syncwarp();
    sts_u32(win_a + ((((pos >> 2) & 32u) ^ 32u) << 2) + 4u * lane_id(), pre);
    const uint32_t nxt = rc_load_half_word(data, len, (pos & ~127u) + 256u);
    __syncwarp();
    return nxt;
}

static __device__ __forceinline__ void rc_row_writeback(const RcBig<8> *cache, uint32_t *rows, int slot, int tag)
{
    const int lane = (int)lane_id();
    const uint4 *s4 = reinterpret_cast<const uint4 *>(&cache[slot]);
    uint4 *g4 = reinterpret_cast<uint4 *>(rows + (size_t)tag * RC_ROW_STRIDE);
#pragma unroll
    for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; if (i < RC_ROW_WORDS / 4) g4[i] = s4[i]; }
}
// evict the least recently used row, load row cxi (or build it: rows not touched since the last renewI are all ones)
static __device__ __noinline__ int rc_row_miss(RcBig<8> *cache, uint32_t *rows, uint32_t gen, int my_tag, uint32_t my_age, int cxi)
{
    const int lane = (int)lane_id();
    const uint32_t key = lane < RC_CACHE_ROWS ? ((my_age << 4) | (uint32_t)lane) : 0xFFFFFFFFu;
    const int slot = (int)(__reduce_min_sync(FULLMASK, key) & 15u);
    const int old = __shfl_sync(FULLMASK, my_tag, slot);
    if (old >= 0) rc_row_writeback(cache, rows, slot, old);
    __syncwarp();
    RcBig<8> &t = cache[slot];
    uint4 *s4 = reinterpret_cast<uint4 *>(&t);
    const uint32_t *grow = rows + (size_t)cxi * RC_ROW_STRIDE;
    const uint4 *g4 = reinterpret_cast<const uint4 *>(grow);
    uint4 v[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; v[k] = i < RC_ROW_WORDS / 4 ? g4[i] : make_uint4(0, 0, 0, 0); }
    const bool fresh = grow[32 * 8 + 33] != gen;
    uint32_t total = 256;
    if (fresh) {
#pragma unroll
        for (int q = 0; q < 8; q++) t.lp[lane * 8 + q] = q + 1;
        t.base[lane] = 8 * lane;
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) { const int i = lane + 32 * k; if (i < RC_ROW_WORDS / 4) s4[i] = v[k]; }
        total = grow[32 * 8 + 32];
    }
    __syncwarp();
    rc_refill_fut(t.hdr, t.fut, total, gen, RC_CLR_STEP);
    return slot;
}

struct RcCoder {
    static constexpr bool kUnrollChannels = true;            // three inlined copies of the colour decoder (sp2_decode.cu)
    static constexpr bool kCanDecodeBool = false;              // EntroCoders.hx:178
    RcShared *shm;                                             // generic pointer (cold paths)
    RcSmall *sm;
    RcBig<8> *cache;
    uint32_t sm_a;                                             // shared-window address of the RcShared; everything else is sm_a + a
                                                               // compile-time offset
    uint32_t lane4;                                            // 4 * lane
    static constexpr uint32_t kCacheOff = (uint32_t)offsetof(RcShared, cache), kWinOff = (uint32_t)offsetof(RcShared, win),
                              kSmallOff = (uint32_t)offsetof(RcShared, small);
    int my_tag;                                                // lane < RC_CACHE_ROWS: context index held by slot `lane`, -1 = empty (and in lanes >= RC_CACHE_ROWS)
    uint32_t my_age, tick;
    uint32_t *rows;
    uint32_t gen;
    uint32_t range, code;
    const uint8_t *data;
    uint32_t len, pos;                                         // pos = index of the next byte
    uint32_t lim;                                              // min(last byte of the window half `pos` is in, len): a symbol that starts beyond it calls catch_up()
    uint32_t w0, w1;                                           // the aligned little-endian word that holds byte `pos`, and the next one
    uint32_t pre;                                              // this lane's word of the window half after next (prefetched from HBM)
    uint32_t nsym;
    bool fail;

    __device__ __forceinline__ bool failed() const { return fail; }
    // A failed frame decodes nothing more (the models stay as they were), at no cost to the symbol chain: failing zeroes
    // `range`, a zero range gives r = 0, every product is then 0 <= code and every later search runs off its table.
    __device__ __forceinline__ void fail_frame() { fail = true; range = 0; }

    // ---- bitstream: a 256-byte circular window in shared memory (word i of the frame lives in slot i & 63), always valid
    //      for at least 128 bytes past `pos`; bytes past the end of the frame read as 0 ----
    __device__ __forceinline__ uint32_t win_word(uint32_t byte_pos) const { return lds_u32(sm_a + kWinOff + (byte_pos & 252u)); }
    // the next (up to 3) bytes, big-endian in the top bytes of the result (byte `pos` in bits 24-31)
    __device__ __forceinline__ uint32_t peek() const { return __byte_perm(__funnelshift_r(w0, w1, (pos & 3u) * 8u), 0u, 0x0123); }
    __device__ __forceinline__ void advance(uint32_t np)
    {
        pos = np;
        w0 = win_word(np); w1 = win_word(np + 4u);
    }
    // `pos` may run past `lim` (into the next window half, or past the end of the data) by the 3 bytes of one symbol; the
    // next symbol notices (its one branch) and calls this before it decodes.
    __device__ __forceinline__ void set_lim() { lim = min(pos | 127u, len); }
    __device__ __forceinline__ void catch_up()
    {
        if (pos > len) range = 0;          // a byte past the end was consumed: in JavaScript `code` is NaN from here on; the symbol
                                           // that took it stands, every later one fails (zero range, see fail_frame)
        if (pos > (lim | 127u)) pre = rc_window_refill(sm_a + kWinOff, data, len, pos, pre);
        set_lim();
    }
    __device__ __forceinline__ void decodeBegin(const uint8_t *src, uint32_t n, uint32_t pos0)   // RangeCoder.hx:19-34
    {
        data = src; len = n; range = 0xFFFFFFFFu;
        pos = pos0 + 1;
        const uint32_t base = pos & ~127u;
        __syncwarp();
        sts_u32(sm_a + kWinOff + (((base >> 2) & 32u) << 2) + lane4, rc_load_half_word(data, len, base));
        sts_u32(sm_a + kWinOff + ((((base + 128u) >> 2) & 32u) << 2) + lane4, rc_load_half_word(data, len, base + 128u));
        pre = rc_load_half_word(data, len, base + 256u);
        __syncwarp();
        w0 = win_word(pos); w1 = win_word(pos + 4u);
        set_lim();
        const uint32_t hi = peek() >> 16;
        advance(pos + 2u);
        if (pos > lim) catch_up();
        code = (hi << 16) | (peek() >> 16);
        advance(pos + 2u);
        if (pos > lim) catch_up();
    }

    // ---- model reset ----
    template <int K>
    __device__ __forceinline__ void init_big(RcBig<K> &t, uint32_t step)
    {
        const int lane = (int)lane_id();
#pragma unroll
        for (int q = 0; q < K; q++) t.lp[lane * K + q] = q + 1;
        t.base[lane] = K * lane;
        __syncwarp();
        rc_refill_fut(t.hdr, t.fut, 32 * K, 0u, step);
    }
    template <int N>
    __device__ __forceinline__ void init_tiny(RcTiny &t, uint32_t step)
    {
        t.P[lane_id()] = lane_id() + 1u;
        __syncwarp();
        rc_refill_tiny(t.P, (uint32_t)N, step);
    }
    // EntroCoders.hx:81-130.  p_only = where the tables only P frames use live: shared memory (P-frame kernel) or the stream's
    // state in HBM (I-frame kernel, which does not stage them)
    __device__ __forceinline__ void renewI(RcSmall *p_only)
    {
        gen = gen + 1;
        for (int t = 0; t < 6; t++) { init_big(sm->ntab[t], 400u); init_tiny<6>(sm->ptypetab[t], 1000u); }
        init_big(p_only->xxtab, 1u); init_big(p_only->ntab2, 20u);
        init_big(p_only->mvtab[0], 100u); init_big(p_only->mvtab[1], 100u);
        for (int t = 0; t < 4; t++) init_tiny<16>(p_only->sxytab[t], 100u);
        init_tiny<5>(p_only->bttab, 10u);
        __syncwarp();
    }

    // I-frame kernel: the tables only P frames use are reset where they live, in the stream's state in HBM
    __device__ __forceinline__ void begin_iframe(const SpJob &J) { renewI(&reinterpret_cast<RcState *>(J.state)->small); }

    // number of renormalisation bytes: `while (range < TOP) range <<= 8` runs once per leading zero byte (3 compares beat clz)
    static __device__ __forceinline__ uint32_t renorm_bytes(uint32_t width)
    {
        return (width < 0x1000000u ? 1u : 0u) + (width < 0x10000u ? 1u : 0u) + (width < 0x100u ? 1u : 0u);
    }

    // RangeCoder.hx:51-80 for tables of N <= 16 symbols, one cumulative count per lane.  ta = shared address of the RcTiny.
    // Searches use REDUX (warp max / min into a uniform register, 18 cycles) instead of ballot + popc + shuffle (22 + 17 +
    // 36, profiles/r02_microbench_latencies.txt): the products ascend with the lane, so the largest one <= code is the
    // interval's low end, the smallest one above it the high end, and the first lane above it the symbol.
    template <int N>
    __device__ __forceinline__ int decode_tiny(uint32_t ta, uint32_t step)
    {
        const int lane = (int)lane_id();
        const uint32_t p = lds_u32(ta + lane4);
        const uint4 h = lds_v4(ta + 28 * 4);                                     // -, cnt, total, inv
        const uint32_t ncnt = h.y + 1u, ntot = h.z + step;
        const uint32_t fnx = lds_u32(ta + 128 + ((ncnt & 31u) << 2));            // reciprocal of total + step (if the batch holds it)
        // the one branch: everything rare is known before the search (the bitstream conditions come from the PREVIOUS symbol)
        if (__builtin_expect((pos > lim) | (ncnt >= 32u) | (ntot > RC_BOT), 0)) return decode_tiny_slow<N>(ta, step);
        nsym++;
        const uint32_t r = udiv1(range, h.z, h.w);
        const uint32_t pr = p * r;
        const bool le = lane < N && pr <= code;
        const uint32_t lo = __reduce_max_sync(FULLMASK, le ? pr : 0u);
        const uint32_t hi = __reduce_min_sync(FULLMASK, (le || lane >= N) ? 0xFFFFFFFFu : pr);
        const uint32_t s = __reduce_min_sync(FULLMASK, le ? 32u : (uint32_t)lane);   // first lane above the value (>= N: off the table)
        const bool good = s < (uint32_t)N;                                       // a search that ran off the table touches nothing
        const uint32_t width = hi - lo, nb = renorm_bytes(width);
        const uint32_t ncode = __funnelshift_l(peek(), code - lo, nb * 8u);      // RangeCoder.hx:36-43
        code = good ? ncode : code;
        range = good ? width << (nb * 8u) : 0u;
        fail = fail | !good;
        advance(good ? pos + nb : pos);
        sts_u32_if(good && !le && lane < 28, ta + lane4, p + step);              // a lane reads back only its own count ...
        if (good) sts_v4(ta + 28 * 4, 0u, ncnt, ntot, fnx);                      // ... and every lane stores the same header
        return good ? (int)s : N - 1;
    }
    template <int N>
    __device__ __forceinline__ int decode_tiny_slow(uint32_t ta, uint32_t step)
    {
        const int lane = (int)lane_id();
        if (pos > lim) catch_up();
        nsym++;
        uint32_t *P = reinterpret_cast<uint32_t *>(shm) + ((ta - sm_a) >> 2);
        const uint32_t p = P[lane], cnt = P[29], tot = P[30], inv = P[31];
        const uint32_t r = udiv1(range, tot, inv);
        const uint32_t pr = p * r;
        const int s = __popc(__ballot_sync(FULLMASK, lane < N && pr <= code));
        if (s >= N) { range = 0; fail = true; return N - 1; }
        const uint32_t below = __shfl_sync(FULLMASK, pr, (s + 31) & 31), hi = __shfl_sync(FULLMASK, pr, s);
        const uint32_t lo = s ? below : 0u, width = hi - lo, nb = renorm_bytes(width);
        code = __funnelshift_l(peek(), code - lo, nb * 8u);
        range = width << (nb * 8u);
        advance(pos + nb);
        const uint32_t ntot = tot + step, ncnt = cnt + 1u;
        const uint32_t np2 = lane >= s ? p + step : p;
        if (ntot > RC_BOT) rc_rescale_tiny<N>(P, np2, step);                     // :70-77
        else {
            if (lane >= s && lane < 28) P[lane] = np2;
            __syncwarp();
            if (ncnt >= 32u) rc_refill_tiny(P, ntot, step);
            else { if (lane == 0) { P[29] = ncnt; P[30] = ntot; P[31] = P[32 + (ncnt & 31u)]; } __syncwarp(); }
        }
        return s;
    }

    // RangeCoder.hx:51-80 (256 / 512 symbols) and :82-130 (colour rows; the reference's 16 group sums are derived data and
    // are not kept) on a table in shared memory: 32 lane bases, then the K prefixes of the owning lane.  ta = shared address.
    // cxi >= 0: a colour row that is NOT in the cache (ta is then meaningless; the slow path loads the row first)
    template <int K>
    __device__ __forceinline__ int decode_big(uint32_t ta, uint32_t step, int miss_cxi = -1)
    {
        const int lane = (int)lane_id();
        const uint32_t base = lds_u32(ta + 32 * K * 4 + lane4);
        const uint4 h = lds_v4(ta + (32 * K + 32) * 4);                          // total, tag, inv, cnt
        const uint32_t ncnt = h.w + 1u, ntot = h.x + step;
        const uint32_t fnx = lds_u32(ta + (32 * K + 36) * 4 + ((ncnt & 31u) << 2));
        if (__builtin_expect((pos > lim) | (ncnt >= 32u) | (ntot > RC_BOT) | (miss_cxi >= 0), 0))
            return decode_big_slow<K>(miss_cxi >= 0 ? row_addr(row_miss(miss_cxi)) : ta, step);
        nsym++;
        const uint32_t r = udiv1(range, h.x, h.z);
        const uint32_t br = base * r;
        const bool le1 = br <= code;                                             // lane 0 has base 0: always true there
        const uint32_t L = __reduce_max_sync(FULLMASK, le1 ? (uint32_t)lane : 0u);
        const uint32_t brL = __reduce_max_sync(FULLMASK, le1 ? br : 0u);
        const uint32_t xa = ta + L * (K * 4) + (lane4 & (K * 4 - 1));            // lane q < K: prefix q of lane L's group
        const uint32_t x = lds_u32(xa);
        const uint32_t px = x * r;
        const bool in = lane < K, le2 = in && px <= code - brL;
        const uint32_t lo = __reduce_max_sync(FULLMASK, le2 ? px : 0u);
        const uint32_t hi = __reduce_min_sync(FULLMASK, (le2 || !in) ? 0xFFFFFFFFu : px);
        const uint32_t m = __reduce_min_sync(FULLMASK, (le2 || !in) ? 32u : (uint32_t)lane);
        // value >= total (not a valid stream)  <=>  L = 31 and every prefix of the last group fits  <=>  no lane above: m = 32
        const bool good = m < (uint32_t)K;
        const uint32_t width = hi - lo, nb = renorm_bytes(width);
        const uint32_t ncode = __funnelshift_l(peek(), code - brL - lo, nb * 8u);
        code = good ? ncode : code;
        range = good ? width << (nb * 8u) : 0u;
        fail = fail | !good;
        advance(good ? pos + nb : pos);
        sts_u32_if(good && in && !le2, xa, x + step);                            // lanes >= K hold copies (q = lane mod K) and store nothing
        sts_u32_if(good && (uint32_t)lane > L, ta + 32 * K * 4 + lane4, base + step);   // a lane reads back only its own base
        if (good) sts_v4(ta + (32 * K + 32) * 4, ntot, h.y, fnx, ncnt);          // every lane: same values
        return good ? (int)(L * K + m) : 32 * K - 1;
    }
    template <int K>
    __device__ __forceinline__ int decode_big_slow(uint32_t ta, uint32_t step)
    {
        const int lane = (int)lane_id();
        if (pos > lim) catch_up();
        nsym++;
        uint32_t *tab = reinterpret_cast<uint32_t *>(shm) + ((ta - sm_a) >> 2);
        const uint32_t base = tab[32 * K + lane], tot = tab[32 * K + 32], tag = tab[32 * K + 33], inv = tab[32 * K + 34], cnt = tab[32 * K + 35];
        const uint32_t r = udiv1(range, tot, inv);
        const uint32_t br = base * r;
        const int L = __popc(__ballot_sync(FULLMASK, br <= code)) - 1;
        const uint32_t x = tab[L * K + (lane & (K - 1))];
        const uint32_t brL = __shfl_sync(FULLMASK, br, L);
        const uint32_t px = x * r;
        const int m = __popc(__ballot_sync(FULLMASK, px <= code - brL) & ((1u << K) - 1u));
        if (m >= K) { range = 0; fail = true; return 32 * K - 1; }
        const uint32_t below = __shfl_sync(FULLMASK, px, (m + 31) & 31), hi = __shfl_sync(FULLMASK, px, m);
        const uint32_t lo = m ? below : 0u, width = hi - lo, nb = renorm_bytes(width);
        code = __funnelshift_l(peek(), code - brL - lo, nb * 8u);
        range = width << (nb * 8u);
        advance(pos + nb);
        const uint32_t ntot = tot + step, ncnt = cnt + 1u;
        if (ntot > RC_BOT) rc_rescale_big<K>(tab, L, m, step);                   // :70-77 / :113-127
        else {
            if (lane < K && lane >= m) tab[L * K + lane] = x + step;
            if (lane > L) tab[32 * K + lane] = base + step;
            __syncwarp();
            if (ncnt >= 32u) rc_refill_fut(tab + 32 * K + 32, tab + 32 * K + 36, ntot, tag, step);
            else { if (lane == 0) { tab[32 * K + 32] = ntot; tab[32 * K + 34] = tab[32 * K + 36 + (ncnt & 31u)]; tab[32 * K + 35] = ncnt; } __syncwarp(); }
        }
        return L * K + m;
    }

    // ---- colour-row cache: 12 fully associative LRU slots, rows decoded in shared memory (a global store invalidates the
    //      L1 line it hits and every symbol updates its row, so decoding rows in place pays an L2 round trip per symbol) ----
    // slot of colour row cxi, or >= RC_CACHE_ROWS when it is not cached (no branch here: decodeClr's one branch covers it)
    __device__ __forceinline__ uint32_t row_lookup(int cxi)
    {
        const uint32_t slot = __reduce_min_sync(FULLMASK, my_tag == cxi ? lane_id() : 255u);
        tick++;
        my_age = lane_id() == slot ? tick : my_age;
        return slot;
    }
    __device__ __forceinline__ uint32_t row_addr(uint32_t slot) const { return sm_a + kCacheOff + slot * (uint32_t)sizeof(RcBig<8>); }
    __device__ __forceinline__ uint32_t row_miss(int cxi)
    {
        const int slot = rc_row_miss(cache, rows, gen, my_tag, my_age, cxi);
        if ((int)lane_id() == slot) { my_tag = cxi; my_age = tick; }
        return (uint32_t)slot;
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

    __device__ __forceinline__ int decodeClr(int cxi)                            // DecodeValUni, RangeCoder.hx:82-130, on a cached colour row
    {
        const uint32_t slot = row_lookup(cxi);
        const bool miss = slot >= (uint32_t)RC_CACHE_ROWS;
        return decode_big<8>(row_addr(miss ? 0u : slot), RC_CLR_STEP, miss ? cxi : -1);
    }
    __device__ __forceinline__ int decodeN(int ptype) { return decode_big<8>(sm_a + kSmallOff + (uint32_t)offsetof(RcSmall, ntab) + (uint32_t)ptype * (uint32_t)sizeof(RcBig<8>), 400u); }   // EntroCoders.hx:142-144
    __device__ __forceinline__ int decodeP(int ptype) { return decode_tiny<6>(sm_a + kSmallOff + (uint32_t)offsetof(RcSmall, ptypetab) + (uint32_t)ptype * (uint32_t)sizeof(RcTiny), 1000u); }
    __device__ __forceinline__ int decodeX() { return decode_big<8>(sm_a + kSmallOff + (uint32_t)offsetof(RcSmall, xxtab), 1u); }
    __device__ __forceinline__ int decodeBT() { return decode_tiny<5>(sm_a + kSmallOff + (uint32_t)offsetof(RcSmall, bttab), 10u); }
    __device__ __forceinline__ int decodeBN() { return decode_big<8>(sm_a + kSmallOff + (uint32_t)offsetof(RcSmall, ntab2), 20u); }
    __device__ __forceinline__ int decodeSXY(int n) { return decode_tiny<16>(sm_a + kSmallOff + (uint32_t)offsetof(RcSmall, sxytab) + (uint32_t)n * (uint32_t)sizeof(RcTiny), 100u); }
    __device__ __forceinline__ int decodeMX() { return decode_big<16>(sm_a + kSmallOff + (uint32_t)offsetof(RcSmall, mvtab), 100u); }
    __device__ __forceinline__ int decodeMY() { return decode_big<16>(sm_a + kSmallOff + (uint32_t)offsetof(RcSmall, mvtab) + (uint32_t)sizeof(RcBig<16>), 100u); }
    __device__ __forceinline__ bool decodeBool() { return false; }

    // ---- per-frame set-up / tear-down: `bytes` of the small tables (all of them, or the part I frames use) travel between
    //      the stream's state in HBM and shared memory ----
    __device__ __forceinline__ void open(const SpJob &J, RcShared *shared, uint32_t bytes)
    {
        RcState *st = reinterpret_cast<RcState *>(J.state);
        shm = shared; sm = &shared->small; cache = shared->cache;
        sm_a = smem_addr(shared);
        asm volatile("mov.u32 %0, %0;" : "+r"(sm_a));        // opaque: or the compiler re-derives the address (S2R + LEA, ~25 cycles) at every use
        lane4 = 4u * lane_id();
        my_tag = -1; my_age = 0; tick = 0;
        rows = st->rows; gen = st->gen;
        fail = false; range = 0; code = 0; data = J.src; len = J.len; pos = 0; lim = 0; w0 = 0; w1 = 0; pre = 0; nsym = 0;
        const uint4 *g = reinterpret_cast<const uint4 *>(&st->small);
        uint4 *s = reinterpret_cast<uint4 *>(&shared->small);
        for (int i = (int)lane_id(); i < (int)(bytes / 16); i += 32) s[i] = g[i];
        __syncwarp();
    }
    __device__ __forceinline__ void close(const SpJob &J, uint32_t bytes)
    {
        RcState *st = reinterpret_cast<RcState *>(J.state);
        flush_rows();
        uint4 *g = reinterpret_cast<uint4 *>(&st->small);
        const uint4 *s = reinterpret_cast<const uint4 *>(sm);
        for (int i = (int)lane_id(); i < (int)(bytes / 16); i += 32) g[i] = s[i];
        if (lane_id() == 0) st->gen = gen;
    }
    // the interface the kernels of sp2_decode.cu use for both coders
    RcShared *bound; uint32_t bound_bytes;
    __device__ __forceinline__ void bind(RcShared *shared, uint32_t small_bytes) { bound = shared; bound_bytes = small_bytes; }
    __device__ __forceinline__ void open_iframe(const SpJob &J) { open(J, bound, bound_bytes); }
    __device__ __forceinline__ void open_pframe(const SpJob &J) { open(J, bound, bound_bytes); }
    __device__ __forceinline__ void close_frame(const SpJob &J) { close(J, bound_bytes); }
    __device__ __forceinline__ void renewI() { renewI(sm); }          // P-frame kernel: every table is in shared memory
};

}  // namespace g2
}  // namespace jsp
