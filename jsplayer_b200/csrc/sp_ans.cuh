// sp_ans.cuh -- ScreenPressor v3 / v4 entropy decode on sm_100a: byte-wise rANS with adaptive context models.
// Replaces reference src/ANS.hx (whole file) and EntroCoderANS (src/EntroCoders.hx:182-313), one stream per warp.
//
//  * rANS state, read position and symbol counter are replicated in every lane; the bitstream is read through a
//    128-byte shared-memory window refilled by one coalesced warp load;
//  * the fixed-size adaptive tables (FixedSizeRansCtx, ANS.hx:54-145: ntab 6x256, ptypetab 6x6, xxtab, ntab2, bttab,
//    sxytab 4x16, mvtab 2x512 -- 21.5 KB) live in shared memory for the whole frame; a lane owns 8 (16) consecutive
//    symbols in registers, the symbol search is one __ballot_sync over "cumFreq of the next symbol > f" and the
//    periodic rebuild (:89-102) is a warp-shuffle prefix sum -- no linear scans;
//  * a colour context (ANS.hx:785-860) is a fixed 64-byte header + 1536-byte body slab in HBM/L2 (12288 per stream).
//    The context kinds escalate exactly as in the reference: raw symbol lists (Cx1/2/3) -> small sorted tables with
//    implicit escapes (Cx4/Cx5) -> Cx6 (<= 40 symbols, move-to-front by count) -> full 256-symbol table (Cx7).
//    One coalesced warp load brings header + the first 512 body bytes (everything but a Cx7); Cx7 runs on the
//    register path of the fixed tables; the small, branchy kinds are bookkept by lane 0 on the shared-memory copy
//    ("one lane does the bookkeeping"), the warp writes the slab back coalesced;
//  * renewI (EntroCoders.hx:216-227) is O(1) for the 12288 contexts: a generation number in the header.
//
// JavaScript typed-array semantics (Uint16Array wrap of freqs / cnts, Uint8Array decTable and symbols) are explicit.
#pragma once
#include "sp_common.cuh"
#include <cstddef>
#include <cstring>
#include <type_traits>

namespace jsp {

constexpr int ANS_SCALE = 4096;                    // Rans.PROB_SCALE
constexpr uint32_t ANS_L = 1u << 23;               // RANS_BYTE_L, ANS.hx:33
constexpr int ANS_B = 131072;                      // Rans.B, ANS.hx:10
constexpr int ANS_NCTX = 3 * 4096;
constexpr int ANS_HDR_BYTES = 64, ANS_BODY_BYTES = 1536;
enum { CXK_NONE = 0, CXK_1, CXK_2, CXK_3, CXK_4, CXK_5, CXK_6, CXK_7 };

// Shared memory on the symbol chain is addressed by 32-bit shared-window addresses kept in registers.  A C++ access through a
// pointer the compiler can trace back to a __shared__ symbol re-derives the window address at the point of use -- S2R
// SR_CgaCtaId + LEA, ~30 cycles of latency in front of the load; the rANS I-frame kernel had 197 of them -- so the hot paths
// use explicit ld.shared / st.shared on an address made opaque once per frame.  The asm statements keep program order.
__device__ __forceinline__ uint32_t a_lds8(uint32_t a)  { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t a_lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t a_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint2 a_lds64(uint32_t a)    { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint4 a_lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void a_sts8(uint32_t a, uint32_t v)  { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void a_sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void a_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t a_opaque_smem(const void *p)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("mov.u32 %0, %0;" : "+r"(a));             // opaque: the compiler must not re-derive it from the symbol
    return a;
}

// ---- fixed-size adaptive table: separate arrays instead of the reference's (freq, cumFreq) pairs ----
// cum[] carries 8 sentinel entries (0xFFFF: above every slot value) behind the table, so a forward scan needs no bound check.
// Tables of 256 / 512 symbols also carry `lut`: the symbol that holds slot 16 * b, for b = 0..255 -- a finer decTable
// (ANS.hx:105-126 looks up decTable[f >> 7] and scans on; this one starts at most 15 symbols short).  Frequencies change
// only in the periodic rebuild (ANS.hx:89-102), which also rewrites lut.
template <int N>
struct FxTab {
    static constexpr int NP = N < 32 ? 32 : N;
    static constexpr int K = NP / 32;
    static constexpr int NLUT = N >= 256 ? 256 : 0;
    typedef typename std::conditional<(N > 256), uint16_t, uint8_t>::type lut_t;
    alignas(16) uint16_t cum[NP + 8];
    alignas(16) uint16_t fr[NP];
    alignas(16) uint16_t cnt[NP];
    uint8_t dec[32];
    uint32_t cntsum;
    uint32_t pad[3];
    alignas(16) lut_t lut[NLUT ? NLUT : 16];
};

// The tables a coded I frame uses come first: the I-frame kernel keeps only that prefix in shared memory (7 CTAs per SM).
struct AnsSmall {
    FxTab<256> ntab[6];
    FxTab<6> ptypetab[6];
    // ---- P frames only ----
    FxTab<256> xxtab, ntab2;
    FxTab<512> mvtab[2];
    FxTab<16> sxytab[4];
    FxTab<5> bttab;
};
constexpr uint32_t ANS_SMALL_I_BYTES = (uint32_t)offsetof(AnsSmall, xxtab);
static_assert(ANS_SMALL_I_BYTES % 16 == 0 && sizeof(AnsSmall) % 16 == 0, "tables are copied in 16-byte units");

struct alignas(16) CxHdr {
    uint32_t gen;          // generation the slab belongs to; any other value reads as "no context yet"
    uint16_t d;            // symbols met
    uint8_t kind;
    uint8_t maxpos;        // SmallContext.maxpos
    uint8_t fshift;        // Cx6.fshift
    uint8_t S;             // SmallContext.S (4 / 16) or Cx6 slots (32 / 64)
    uint16_t pad0;
    uint32_t cntsum;       // Cx5.cntsum | Cx6 cnts[S] (Uint16) | Cx7 cntsum
    uint8_t dec[32];       // Cx7 decTable
    uint8_t pad1[16];
};
static_assert(sizeof(CxHdr) == ANS_HDR_BYTES, "CxHdr layout");

// body layouts (byte offsets): kinds 1-3 symb[256] @0 | kinds 4/5 symbols[16] @0, freqs u16[16] @16 |
// kind 6 symbols[64] @0, freq u16[64] @64, cumFreq u16[64] @192, cnts u16[64] @320 |
// kind 7 cumFreq u16[256] @0, freq u16[256] @512, cnts u16[256] @1024
constexpr int B_SC_FR = 16, B6_FR = 64, B6_CUM = 192, B6_CNT = 320, B6_BYTES = 448, B7_FR = 512, B7_CNT = 1024;

struct AnsState {                                  // per stream, in HBM
    AnsSmall small;
    uint32_t gen;
    uint32_t pad[3];
    uint4 *hdrs;                                   // ANS_NCTX * 4 uint4
    uint4 *bodies;                                 // ANS_NCTX * 96 uint4
};

// A cached colour context of a small kind (None, Cx1..Cx6): header + the first 512 body bytes.  Global stores
// invalidate the L1 lines they hit and every symbol updates its context, so a context read straight from global memory
// costs an L2 round trip per symbol; screen content keeps returning to a handful of contexts (12 LRU slots catch
// 93-98 %).  Cx7 contexts (1.5 KB) are rare there and stay in global memory.
constexpr int ANS_CACHE_SLOTS = 12;
struct alignas(16) AnsSlot {
    CxHdr hdr;
    uint8_t body[512];
};

constexpr uint32_t ANS_WIN = 256;                  // bitstream window (bytes)
struct AnsWork {                                   // everything but the small tables
    AnsSlot cache[ANS_CACHE_SLOTS];
    alignas(16) uint8_t big[ANS_BODY_BYTES];       // Cx7 under construction / Cx6 rescale temporaries
    alignas(16) uint8_t win[ANS_WIN];
    int res_c, res_freq, res_cum, res_wb;          // lane 0 -> warp
};
struct AnsShared {
    AnsSmall small;
    AnsWork work;
};

struct CxRes { int c, freq, cum; };

// ---- register-resident table ops -------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void ld_u16(const uint16_t *p, uint32_t (&v)[K])
{
    if constexpr (K == 1) { v[0] = *p; }
    else {
#pragma unroll
        for (int q = 0; q < K / 8; q++) {
            const uint4 w = reinterpret_cast<const uint4 *>(p)[q];
            v[8 * q + 0] = w.x & 0xFFFFu; v[8 * q + 1] = w.x >> 16; v[8 * q + 2] = w.y & 0xFFFFu; v[8 * q + 3] = w.y >> 16;
            v[8 * q + 4] = w.z & 0xFFFFu; v[8 * q + 5] = w.z >> 16; v[8 * q + 6] = w.w & 0xFFFFu; v[8 * q + 7] = w.w >> 16;
        }
    }
}
template <int K>
__device__ __forceinline__ void st_u16(uint16_t *p, const uint32_t (&v)[K])
{
    if constexpr (K == 1) { *p = (uint16_t)v[0]; }
    else {
#pragma unroll
        for (int q = 0; q < K / 8; q++)
            reinterpret_cast<uint4 *>(p)[q] = make_uint4(v[8 * q] | (v[8 * q + 1] << 16), v[8 * q + 2] | (v[8 * q + 3] << 16),
                                                         v[8 * q + 4] | (v[8 * q + 5] << 16), v[8 * q + 6] | (v[8 * q + 7] << 16));
    }
}

// FixedSizeRansCtx.decode + incrCnt (ANS.hx:85-126) on a table whose lane-owned symbols are in registers.
// dec = the table's decTable (shared memory).  Returns the symbol; freq / cumf = its interval BEFORE the update.
template <int N, int K>
__device__ __forceinline__ int fx_core(uint32_t (&cum)[K], uint32_t (&fr)[K], uint32_t (&cnt)[K], uint8_t *dec,
                                       uint32_t &cntsum, int f, int &freq, int &cumf, bool &rebuilt, int &owner)
{
    const int lane = (int)lane_id(), j0 = lane * K;
    const int c0 = dec[(f >> 7) & 31];
    const uint32_t nxt = __shfl_down_sync(FULLMASK, cum[0], 1);
    uint32_t mask = 0;
#pragma unroll
    for (int q = 0; q < K; q++) {
        const int j = j0 + q;
        const uint32_t cn = q + 1 < K ? cum[(q + 1) % K] : nxt;
        if (j >= c0 && j < N - 1 && (int)cn > f) mask |= 1u << q;
    }
    const uint32_t ball = __ballot_sync(FULLMASK, mask != 0);
    int c = N - 1;
    if (ball) {
        const int lw = __ffs(ball) - 1;
        const uint32_t m = __shfl_sync(FULLMASK, mask, lw);
        c = lw * K + __ffs(m) - 1;
    }
    owner = c / K;
    const int qs = c % K;
    uint32_t sf = 0, sc = 0;
#pragma unroll
    for (int q = 0; q < K; q++) if (q == qs) { sf = fr[q]; sc = cum[q]; }
    freq = (int)__shfl_sync(FULLMASK, sf, owner);
    cumf = (int)__shfl_sync(FULLMASK, sc, owner);
    if (lane == owner) {
#pragma unroll
        for (int q = 0; q < K; q++) if (q == qs) cnt[q] = (cnt[q] + 16u) & 0xFFFFu;
    }
    cntsum += 16;
    rebuilt = cntsum + 16 > (uint32_t)ANS_SCALE;
    if (rebuilt) {                                                   // ANS.hx:89-102
        uint32_t s = 0;
#pragma unroll
        for (int q = 0; q < K; q++) if (j0 + q < N) s += cnt[q];
        uint32_t incl = s;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t o = __shfl_up_sync(FULLMASK, incl, dd); if (lane >= dd) incl += o; }
        uint32_t cf = incl - s, ns = 0;
#pragma unroll
        for (int q = 0; q < K; q++) {
            const int j = j0 + q;
            if (j < N) {
                const uint32_t frq = cnt[q];
                fr[q] = frq; cum[q] = cf & 0xFFFFu;
                const int k0 = (int)((cf + 127u) >> 7), k1 = (((int)(cf + frq) - 1) >> 7) + 1;
                for (int k = k0; k < k1; k++) if (k < 32) dec[k] = (uint8_t)j;
                cf += frq;
                cnt[q] = (cnt[q] - (frq >> 1)) & 0xFFFFu;
                ns += cnt[q];
            }
        }
        cntsum = __reduce_add_sync(FULLMASK, ns);
        __syncwarp();
    }
    return c;
}

// ---- lane-0 bookkeeping of the small colour-context kinds on the shared-memory copy (ANS.hx:155-704) ----
namespace cx {

__device__ __forceinline__ uint16_t *u16p(uint8_t *B, int off) { return reinterpret_cast<uint16_t *>(B + off); }
__device__ __forceinline__ int scale_shift(int tot, int &scaled)      // `while (tot <= PROB_SCALE/2) { tot <<= 1; shift++; }`
{
    int shift = 0;
    while (tot <= ANS_SCALE / 2 && tot > 0) { tot <<= 1; shift++; }
    scaled = tot;
    return shift;
}
static __device__ void insort(uint8_t *a, int n)                             // Sorter.insort, ANS.hx:862-872
{
    for (int i = 1; i < n; i++) {
        int j = i;
        while (j > 0 && a[j - 1] > a[j]) { const uint8_t t = a[j]; a[j] = a[j - 1]; a[j - 1] = t; j--; }
    }
}
static __device__ void fill_dec(uint8_t *dec, int i, int cf, int fr)
{
    const int k0 = (cf + 127) >> 7, k1 = ((cf + fr - 1) >> 7) + 1;
    for (int k = k0; k < k1; k++) if (k >= 0 && k < 32) dec[k] = (uint8_t)i;
}

// SmallContext.create, :226-238 -- the Cx1 list is in B[0..d)
static __device__ void sc_create(CxHdr &H, uint8_t *B, int S, int c)
{
    const int d = H.d;
    uint16_t *fq = u16p(B, B_SC_FR);
    insort(B, d);
    for (int i = d; i < 16; i++) B[i] = 0;
    H.S = (uint8_t)S; H.maxpos = 0;
    for (int i = 0; i < 16; i++) fq[i] = 0;
    for (int i = 0; i < d; i++) {
        if (B[i] == c) { fq[i] = 100; H.maxpos = (uint8_t)i; } else fq[i] = 50;
    }
}
static __device__ void sc_rescale(CxHdr &H, uint8_t *B, int &totFr)          // :254-261
{
    uint16_t *fq = u16p(B, B_SC_FR);
    int s = 256 - H.d;
    for (int i = 0; i < H.d; i++) { fq[i] = (uint16_t)(fq[i] - (fq[i] >> 1)); s += fq[i]; }
    totFr = s;
}
static __device__ bool sc_add(CxHdr &H, uint8_t *B, int pos, int c, int &totFr)   // addSymb, :240-252
{
    if (H.d == H.S) return false;
    uint16_t *fq = u16p(B, B_SC_FR);
    for (int i = H.d - 1; i >= pos; i--) { B[i + 1] = B[i]; fq[i + 1] = fq[i]; }
    B[pos] = (uint8_t)c; fq[pos] = 50; H.d++;
    if (H.maxpos >= pos) H.maxpos++;
    totFr += 50;
    if (totFr + 50 > ANS_SCALE) sc_rescale(H, B, totFr);
    return true;
}
// SmallContext.decodeSC, :263-309
static __device__ bool sc_decode(CxHdr &H, uint8_t *B, int someFreq, CxRes &r, int totFr0, int &totFr)
{
    uint16_t *fq = u16p(B, B_SC_FR);
    totFr = totFr0;
    int tot;
    const int shift = scale_shift(totFr0, tot);
    someFreq >>= shift;
    const int bonus = (ANS_SCALE - tot) >> shift;
    const int mp = H.maxpos, d = H.d;
    const uint16_t maxFreq = fq[mp];
    fq[mp] = (uint16_t)(maxFreq + bonus);
    int cumFr = 0, lastSymb = 0, pos = 0;
    while (pos < d) {
        const int s = B[pos];
        const int startFr = cumFr + s - lastSymb;
        if (someFreq < startFr) {
            r.c = someFreq - cumFr + lastSymb;
            r.cum = someFreq << shift; r.freq = 1 << shift;
            fq[mp] = maxFreq;
            return sc_add(H, B, pos, r.c, totFr);
        }
        const int frq = fq[pos];
        if (startFr + frq > someFreq) {
            r.c = s;
            cumFr += s - lastSymb;
            r.cum = cumFr << shift; r.freq = frq << shift;
            fq[mp] = maxFreq;
            fq[pos] = (uint16_t)(fq[pos] + 50); totFr += 50;
            if (pos != H.maxpos && fq[pos] > fq[H.maxpos]) H.maxpos = (uint8_t)pos;
            if (totFr + 50 > ANS_SCALE) sc_rescale(H, B, totFr);
            return true;
        }
        cumFr += s - lastSymb + frq;
        lastSymb = s + 1;
        pos++;
    }
    fq[mp] = maxFreq;
    r.c = lastSymb + someFreq - cumFr;
    r.cum = someFreq << shift; r.freq = 1 << shift;
    return sc_add(H, B, pos, r.c, totFr);
}
static __device__ void c5_calcsum(CxHdr &H, uint8_t *B)                      // :374-378
{
    const uint16_t *fq = u16p(B, B_SC_FR);
    int t = 256 - H.d;
    for (int i = 0; i < H.d; i++) t += fq[i];
    H.cntsum = (uint32_t)t;
}
// Cx5.createFrom4, :350-372 (in place)
static __device__ void c5_from4(CxHdr &H, uint8_t *B, int c)
{
    uint16_t *fq = u16p(B, B_SC_FR);
    uint8_t os[4]; uint16_t of[4];
    const int dd = H.d;
    for (int i = 0; i < 4; i++) { os[i] = B[i]; of[i] = fq[i]; }
    for (int i = 0; i < 16; i++) { B[i] = 0; fq[i] = 0; }
    H.S = 16; H.maxpos = 0;                                           // a new Cx5: maxpos starts at 0 (:223)
    int i = 0, totFr = 0;
    while (i < dd && os[i] < c) { B[i] = os[i]; fq[i] = of[i]; totFr += of[i]; i++; }
    int j = i;
    B[j] = (uint8_t)c; fq[j] = 50; totFr += 50; j++;
    while (i < dd) { B[j] = os[i]; fq[j] = of[i]; totFr += of[i]; i++; j++; }
    H.d = (uint16_t)(dd + 1);
    if (totFr > ANS_SCALE) { int t; sc_rescale(H, B, t); }
    c5_calcsum(H, B);
    H.kind = CXK_5;
}

// ---- Cx6 ----
static __device__ void c6_init(CxHdr &H, uint8_t *B, int S)
{
    H.S = (uint8_t)S;
    uint32_t *w = reinterpret_cast<uint32_t *>(B);
    for (int i = 0; i < B6_BYTES / 4; i++) w[i] = 0;
    H.cntsum = 0;
}
static __device__ void c6_calcsum(CxHdr &H, uint8_t *B)                      // :571-578
{
    const uint16_t *cn = u16p(B, B6_CNT);
    const int shft = H.fshift > 0 ? H.fshift - 1 : 0;
    int sum = (256 - H.d) << shft;
    for (int i = 0; i < H.S; i++) sum += cn[i];
    H.cntsum = (uint32_t)sum & 0xFFFFu;
}
static __device__ void c6_rescale(CxHdr &H, uint8_t *B, uint8_t *big)        // rescaleDec, :580-604
{
    uint16_t *fr = u16p(B, B6_FR), *cm = u16p(B, B6_CUM), *cn = u16p(B, B6_CNT);
    uint16_t *_cnts = reinterpret_cast<uint16_t *>(big), *_cum = _cnts + 256;
    const int sh = H.fshift > 0 ? H.fshift - 1 : 0;
    const int c0 = 1 << sh, d = H.d;
    for (int i = 0; i < 256; i++) _cnts[i] = (uint16_t)c0;
    for (int i = 0; i < d; i++) _cnts[B[i]] = cn[i];
    int cumFr = 0;
    for (int i = 0; i < 256; i++) { _cum[i] = (uint16_t)cumFr; cumFr += _cnts[i]; }
    if (H.fshift > 0) H.fshift--;
    const int shft = H.fshift > 0 ? H.fshift - 1 : 0;
    int cntsum = (256 - d) << shft;
    for (int i = 0; i < d; i++) {
        cn[i] = (uint16_t)(cn[i] - (cn[i] >> 1));
        cntsum += cn[i];
        const int idx = B[i];
        fr[i] = _cnts[idx]; cm[i] = _cum[idx];
    }
    H.cntsum = (uint32_t)cntsum & 0xFFFFu;
}
static __device__ void c6_incr(CxHdr &H, uint8_t *B, uint8_t *big, int pos)  // incrCntDec, :680-696
{
    uint16_t *fr = u16p(B, B6_FR), *cm = u16p(B, B6_CUM), *cn = u16p(B, B6_CNT);
    const int step = 25 << H.fshift;
    if (pos >= 0) cn[pos] = (uint16_t)(cn[pos] + step);
    H.cntsum = (H.cntsum + (uint32_t)step) & 0xFFFFu;
    if (pos > 0 && cn[pos] > cn[pos - 1]) {
        const uint16_t tc = cn[pos]; cn[pos] = cn[pos - 1]; cn[pos - 1] = tc;
        const uint16_t tf = fr[pos]; fr[pos] = fr[pos - 1]; fr[pos - 1] = tf;
        const uint16_t tm = cm[pos]; cm[pos] = cm[pos - 1]; cm[pos - 1] = tm;
        const uint8_t ts = B[pos]; B[pos] = B[pos - 1]; B[pos - 1] = ts;
    }
    if ((int)H.cntsum + step > ANS_SCALE) c6_rescale(H, B, big);
}
// Cx6.createFrom5, :431-505 (in place: B holds the Cx5; c did not fit)
static __device__ void c6_from5(CxHdr &H, uint8_t *B, uint8_t *big, int c)
{
    uint8_t *os = big + 1024; uint16_t *of = reinterpret_cast<uint16_t *>(big + 1040);
    const int oldd = H.d;
    for (int i = 0; i < 16; i++) { os[i] = B[i]; of[i] = u16p(B, B_SC_FR)[i]; }
    c6_init(H, B, 32);
    uint16_t *fr = u16p(B, B6_FR), *cm = u16p(B, B6_CUM), *cn = u16p(B, B6_CNT);
    int totFr = 256 - oldd;
    for (int i = 0; i < oldd; i++) totFr += of[i];
    int tot;
    const int shift = scale_shift(totFr, tot);
    int cumFr = 0, lastSymb = 0;
    for (int pos = 0; pos < oldd; pos++) {
        const int s = os[pos];
        cumFr += s - lastSymb;
        const int cfr = of[pos], f = cfr << shift;
        fr[pos] = (uint16_t)f; cm[pos] = (uint16_t)(cumFr << shift);
        cn[pos] = (uint16_t)(f - (f >> 1));
        B[pos] = (uint8_t)s;
        cumFr += cfr;
        lastSymb = s + 1;
    }
    H.fshift = (uint8_t)shift;
    const int fr_freq = 1 << shift; int fr_cum = 0;
    if (c > 0) {
        int lowerSym = -1, lfreq = 0, lcum = 0;
        for (int i = 0; i < oldd; i++) {
            const int s = B[i];
            if (s > lowerSym && s < c) { lowerSym = s; lfreq = fr[i]; lcum = cm[i]; }
        }
        fr_cum = lfreq > 0 ? lcum + lfreq + ((c - lowerSym - 1) << shift) : (c << shift);
    }
    fr[oldd] = (uint16_t)fr_freq; cm[oldd] = (uint16_t)fr_cum;
    cn[oldd] = (uint16_t)(fr_freq - (fr_freq >> 1));
    B[oldd] = (uint8_t)c;
    H.d = (uint16_t)(oldd + 1);
    const int step = 25 << shift;
    cn[oldd] = (uint16_t)(cn[oldd] + step);
    H.cntsum = (H.cntsum + (uint32_t)step) & 0xFFFFu;
    if ((int)H.cntsum + step > ANS_SCALE) c6_rescale(H, B, big);
    c6_calcsum(H, B);
    const int d = H.d;
    for (int i = 0; i < d - 1; i++)                                   // sort by freqs, descending
        for (int j = i + 1; j < d; j++) {
            const uint16_t fj = fr[j], fi = fr[i];
            if (fj > fi) {
                const uint16_t cfi = cm[i], cfj = cm[j];
                fr[i] = fj; cm[i] = cfj; fr[j] = fi; cm[j] = cfi;
                const uint16_t tc = cn[i]; cn[i] = cn[j]; cn[j] = tc;
                const uint8_t ts = B[i]; B[i] = B[j]; B[j] = ts;
            }
        }
    H.kind = CXK_6;
}
// Cx6.createFrom2, :507-555 (B holds the Cx2 list; c was met the second time)
static __device__ void c6_from2(CxHdr &H, uint8_t *B, uint8_t *big, int c, int f0)
{
    const int oldd = H.d;
    uint8_t *ss = big + 1024;
    for (int i = 0; i < oldd; i++) ss[i] = B[i];
    insort(ss, oldd);
    c6_init(H, B, oldd <= 32 ? 32 : 64);
    uint16_t *fr = u16p(B, B6_FR), *cm = u16p(B, B6_CUM), *cn = u16p(B, B6_CNT);
    const int totFr = 256 - oldd + oldd * f0 + f0;
    int tot;
    const int shift = scale_shift(totFr, tot);
    int cumFr = 0, lastSymb = 0, newSymbPos = 0;
    for (int pos = 0; pos < oldd; pos++) {
        const int s = ss[pos];
        cumFr += s - lastSymb;
        int cfr;
        if (s == c) { newSymbPos = pos; cfr = f0 * 2; } else cfr = f0;
        const int f = cfr << shift;
        fr[pos] = (uint16_t)f; cm[pos] = (uint16_t)(cumFr << shift);
        B[pos] = (uint8_t)s;
        cn[pos] = (uint16_t)(f - (f >> 1));
        cumFr += cfr;
        lastSymb = s + 1;
    }
    H.d = (uint16_t)oldd; H.fshift = (uint8_t)shift;
    c6_calcsum(H, B);
    if (newSymbPos > 0) {                                             // put that symbol on the 0th position
        const uint16_t fr0 = fr[0], cf0 = cm[0], frc = fr[newSymbPos], cfc = cm[newSymbPos];
        fr[0] = frc; cm[0] = cfc; fr[newSymbPos] = fr0; cm[newSymbPos] = cf0;
        const uint8_t sym0 = B[0]; const uint16_t cnt0 = cn[0], cntc = cn[newSymbPos];
        cn[0] = cntc; cn[newSymbPos] = cnt0;
        B[0] = (uint8_t)c; B[newSymbPos] = sym0;
    }
    H.kind = CXK_6;
}
static __device__ int c6_add(CxHdr &H, uint8_t *B, int c, int freq, int cum)     // addDec, :652-661
{
    if (H.d >= 40 || H.d >= H.S) return -1;
    const int pos = H.d;
    B[pos] = (uint8_t)c; u16p(B, B6_FR)[pos] = (uint16_t)freq; u16p(B, B6_CUM)[pos] = (uint16_t)cum;
    u16p(B, B6_CNT)[pos] = (uint16_t)(freq - (freq >> 1));
    H.d++;
    return pos;
}
// Cx6.decode, :606-650; false = the context must be upgraded to Cx7 (r.c is the symbol)
static __device__ bool c6_decode(CxHdr &H, uint8_t *B, uint8_t *big, int someFreq, CxRes &r)
{
    const uint16_t *fr = u16p(B, B6_FR), *cm = u16p(B, B6_CUM);
    int lfreq = 0, lcum = 0, lowerSym = 0;
    const int d = H.d;
    for (int i = 0; i < d; i++) {
        const int cf = cm[i];
        if (cf <= someFreq) {
            const int f = fr[i];
            if (cf + f > someFreq) { r.c = B[i]; r.freq = f; r.cum = cf; c6_incr(H, B, big, i); return true; }
            if (cf >= lcum) { lfreq = f; lcum = cf; lowerSym = B[i]; }
        }
    }
    const int fr_freq = 1 << H.fshift; int fr_cum, c;
    if (lfreq > 0) {
        const int cumFr = lcum + lfreq;
        const int xx = (someFreq - cumFr) >> H.fshift;
        c = xx + lowerSym + 1;
        fr_cum = lcum + lfreq + (xx << H.fshift);
    } else { c = someFreq >> H.fshift; fr_cum = c << H.fshift; }
    r.freq = fr_freq; r.cum = fr_cum; r.c = c;
    int p = c6_add(H, B, c, fr_freq, fr_cum);
    if (p < 0) {
        if (H.S == 64) return false;
        H.S = 64;                                                     // growDec, :663-678 (slots 32..63 are already zero)
        p = c6_add(H, B, c, fr_freq, fr_cum);
        if (p < 0) return false;
    }
    c6_incr(H, B, big, p);
    return true;
}

// ---- Cx7 under construction in `big` (cumFreq @0, freq @512, cnts @1024), decTable in the header ----
static __device__ void c7_from3(CxHdr &H, uint8_t *B, uint8_t *big, int c)   // :711-739
{
    uint16_t *cm = reinterpret_cast<uint16_t *>(big), *fr = cm + 256, *cn = cm + 512;
    for (int i = 0; i < 256; i++) { fr[i] = 1; cn[i] = 1; }
    const int d = H.d;
    const int f0 = (ANS_SCALE - (256 - d)) / (d + 1);
    const int c0 = f0 - (f0 >> 1);
    for (int i = 0; i < d; i++) { const int s = B[i]; fr[s] = (uint16_t)f0; cn[s] = (uint16_t)c0; }
    fr[c] = (uint16_t)(fr[c] + f0);
    cn[c] = (uint16_t)(cn[c] + 16);
    for (int k = 0; k < 32; k++) H.dec[k] = 0;
    int cntsum = 0, cf = 0;
    for (int i = 0; i < 256; i++) {
        cntsum += cn[i];
        cm[i] = (uint16_t)cf;
        const int f = fr[i];
        fill_dec(H.dec, i, cf, f);
        cf += f;
    }
    H.cntsum = (uint32_t)cntsum;
    H.kind = CXK_7;
}
static __device__ void c7_from6(CxHdr &H, uint8_t *B, uint8_t *big)          // :741-771
{
    uint16_t *cm = reinterpret_cast<uint16_t *>(big), *fr = cm + 256, *cn = cm + 512;
    const uint16_t *fr6 = u16p(B, B6_FR), *cm6 = u16p(B, B6_CUM), *cn6 = u16p(B, B6_CNT);
    for (int i = 0; i < 256; i++) { cm[i] = 0; fr[i] = 0; cn[i] = 0; }
    for (int k = 0; k < 32; k++) H.dec[k] = 0;
    const int S = H.S;
    for (int i = 0; i < S; i++) if (cn6[i] > 0) {
        const int s = B[i];
        fr[s] = fr6[i]; cm[s] = cm6[i]; cn[s] = cn6[i];
    }
    const int funmet = 1 << H.fshift, cntUnmet = funmet - (funmet >> 1);
    int cumFr = 0;
    for (int i = 0; i < 256; i++) {
        int f;
        if (fr[i] > 0) f = fr[i];
        else { fr[i] = (uint16_t)funmet; cm[i] = (uint16_t)cumFr; cn[i] = (uint16_t)cntUnmet; f = funmet; }
        fill_dec(H.dec, i, cumFr, f);
        cumFr += f;
    }
    // cntsum = c6.cnts[S] is already in the header
    H.kind = CXK_7;
}

enum { FOUND, ADDED, NOROOM };
static __device__ int find_or_add(CxHdr &H, uint8_t *B, int c, int cap)      // SymbList.findOrAdd, :163-171
{
    const int d = H.d;
    for (int i = 0; i < d; i++) if (B[i] == c) return FOUND;
    if (d < cap) { B[d] = (uint8_t)c; H.d++; return ADDED; }
    return NOROOM;
}

// Context.decode for kinds 4-6 (ANS.hx:795-810). Returns the number of body bytes to write back, or -1 when a
// Cx7 was built in `big`.
__device__ __forceinline__ int decode_small(CxHdr &H, uint8_t *B, uint8_t *big, int someFreq, CxRes &r)
{
    int tf;
    switch (H.kind) {
    case CXK_4: {
        const uint16_t *fq = u16p(B, B_SC_FR);
        const int tot = fq[0] + fq[1] + fq[2] + fq[3] + 256 - H.d;    // :320
        if (!sc_decode(H, B, someFreq, r, tot, tf)) { c5_from4(H, B, r.c); }
        return 48;
    }
    case CXK_5: {
        const bool ok = sc_decode(H, B, someFreq, r, (int)H.cntsum, tf);
        H.cntsum = (uint32_t)tf;
        if (!ok) { c6_from5(H, B, big, r.c); return B6_BYTES; }
        return 48;
    }
    default:                                                          // CXK_6
        if (!c6_decode(H, B, big, someFreq, r)) { c7_from6(H, B, big); return -1; }
        return B6_BYTES;
    }
}

// Context.update for kinds None-3 after a raw symbol (ANS.hx:812-859). Same return convention.
__device__ __forceinline__ int update_raw(CxHdr &H, uint8_t *B, uint8_t *big, int c, int f0, uint32_t gen)
{
    int kind = H.gen == gen ? H.kind : CXK_NONE;
    switch (kind) {
    case CXK_NONE:
        H.gen = gen; H.kind = CXK_1; H.d = 1; H.maxpos = 0; H.fshift = 0; H.S = 0; H.cntsum = 0;
        B[0] = (uint8_t)c;
        return 16;
    case CXK_1:
        switch (find_or_add(H, B, c, 14)) {
        case FOUND:
            if (H.d <= 4) { sc_create(H, B, 4, c); H.kind = CXK_4; }
            else { sc_create(H, B, 16, c); c5_calcsum(H, B); H.kind = CXK_5; }     // Cx5.fromCx1, :337-342
            return 48;
        case NOROOM: B[H.d] = (uint8_t)c; H.d++; H.kind = CXK_2; return 16;          // new Cx2(c1, c), :188-197
        default: return 16;
        }
    case CXK_2:
        switch (find_or_add(H, B, c, 64)) {
        case FOUND: c6_from2(H, B, big, c, f0); return B6_BYTES;
        case NOROOM: B[H.d] = (uint8_t)c; H.d++; H.kind = CXK_3; return 80;          // new Cx3(c2, c), :199-208
        default: return 64;
        }
    default:                                                          // CXK_3
        if (find_or_add(H, B, c, 256) == FOUND) { c7_from3(H, B, big, c); return -1; }
        return 256;
    }
}

}  // namespace cx

// The symbol that triggers a table's rebuild (ANS.hx:89-102): warp-parallel register path, a lane owns K consecutive symbols.
// Out of line on purpose (scalar arguments, result in registers: {symbol, freq | cumFreq << 16}): one copy per table size.
template <int N>
static __device__ __noinline__ uint2 fx_rebuild_symbol(FxTab<N> *tp, int f)
{
        FxTab<N> &t = *tp;
        constexpr int K = FxTab<N>::K;
        const int lane = (int)lane_id(), j0 = lane * K;
        uint32_t cum[K], fr[K], cnt[K];
        ld_u16<K>(t.cum + j0, cum); ld_u16<K>(t.fr + j0, fr); ld_u16<K>(t.cnt + j0, cnt);
        uint32_t cs = t.cntsum;
        int freq, cumf, owner; bool rebuilt;
        const int c = fx_core<N, K>(cum, fr, cnt, t.dec, cs, f, freq, cumf, rebuilt, owner);
        __syncwarp();
        if constexpr (K == 1) { if (j0 < N) st_u16<K>(t.cum + j0, cum); }   // entries >= N stay sentinels
        else st_u16<K>(t.cum + j0, cum);
        st_u16<K>(t.fr + j0, fr); st_u16<K>(t.cnt + j0, cnt);
        t.cntsum = cs;                                                 // every lane stores the same value
        if constexpr (FxTab<N>::NLUT != 0) {
            // lut[b] = the symbol holding slot 16 * b: every symbol marks the first bucket start inside its interval, a prefix
            // maximum spreads the marks (symbol indices grow with the slot)
            static_assert(FxTab<N>::NLUT == 256, "8 buckets per lane");
            typedef typename FxTab<N>::lut_t lt;
            lt *lut = t.lut;
#pragma unroll
            for (int i = 0; i < 8; i++) lut[8 * lane + i] = 0;
            __syncwarp();
            if (rebuilt) {
#pragma unroll
                for (int q = 0; q < K; q++) {
                    const uint32_t cf = cum[q], b0 = (cf + 15u) >> 4;
                    if (j0 + q < N && b0 < 256u && (b0 << 4) < cf + fr[q]) lut[b0] = (lt)(j0 + q);
                }
            }
            __syncwarp();
            uint32_t m[8], run = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { run = max(run, (uint32_t)lut[8 * lane + i]); m[i] = run; }
            uint32_t incl = run;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t o = __shfl_up_sync(FULLMASK, incl, dd); if (lane >= dd) incl = max(incl, o); }
            uint32_t excl = __shfl_up_sync(FULLMASK, incl, 1);
            if (lane == 0) excl = 0;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++) lut[8 * lane + i] = (lt)max(m[i], excl);
        }
        __syncwarp();
        return make_uint2((uint32_t)c, (uint32_t)freq | ((uint32_t)cumf << 16));
}


#ifdef JSP_PROFILE_SECTIONS
__device__ unsigned long long g_ans_prof[16];
#define JSP_AT(k) { const long long _n = clock64(); aprof[k] += _n - _at; aprof[8 + (k)]++; _at = _n; }
#else
#define JSP_AT(k)
#endif

struct AnsCoder {
    static constexpr bool kCanDecodeBool = true;                      // EntroCoders.hx:257
    static constexpr bool kUnrollChannels = false;                    // one copy of decodeClr per kernel (sp2_decode.cu)
    AnsSmall *small;                                                  // shared memory (the I-frame kernel maps only the I-frame prefix)
    AnsWork *wk;
    uint32_t small_bytes;                                             // how much of AnsSmall is mapped
    uint32_t sa_small, sa_wk;                                         // the same two as shared-window addresses (see a_lds8)
    uint4 *hdrs, *bodies;
    int my_tag;                                                       // lane < ANS_CACHE_SLOTS: context held by slot `lane`, -1 = empty
    uint32_t my_age, tick;                                            // LRU stamps
    uint32_t gen;
    int f0;
    uint32_t x;                                                       // rANS state (low 32 bits; ANS.hx:6)
    const uint8_t *data;
    uint32_t len, pos, wbase;
    int nDec;
    uint32_t nsym;                                                    // symbols decoded in this frame (reporting only)
#ifdef JSP_PROFILE_SECTIONS
    long long aprof[16];
#endif
    bool overrun, fail;

    __device__ __forceinline__ bool failed() const { return fail; }
    // a failure found by the frame loop.  (rANS arithmetic stays well defined on garbage -- 32-bit state, bounded tables --
    // so this coder simply goes on until the loop's next check, exactly as the oracle does.)
    __device__ __forceinline__ void fail_frame() { fail = true; }

    __device__ __forceinline__ void refill(uint32_t at)              // window <- the ANS_WIN bytes around `at` (one coalesced warp load)
    {
        __syncwarp();
        wbase = at & ~7u;                                             // the window starts (almost) at the read position
        const int lane = (int)lane_id();
        uint32_t w[2] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t p = wbase + 8u * lane + k;
            if (p < len) w[k >> 2] |= (uint32_t)__ldg(data + p) << (8 * (k & 3));
        }
        reinterpret_cast<uint2 *>(wk->win)[lane] = make_uint2(w[0], w[1]);
        __syncwarp();
    }
    __device__ __forceinline__ uint32_t rbyte()                       // data[pos++]; out of bounds reads as 0 after `|`
    {
        uint32_t b = 0;
        if (pos < len) {
            if (pos - wbase >= ANS_WIN) refill(pos);
            b = wk->win[pos - wbase];
        } else overrun = true;
        pos++;
        return b;
    }
    __device__ __forceinline__ void reinit(uint32_t i)                                // Rans.reinitImpl, ANS.hx:22-31
    {
        pos = i;
        uint32_t v = rbyte();
        v |= rbyte() << 8; v |= rbyte() << 16; v |= rbyte() << 24;
        x = v;
    }
    __device__ __forceinline__ void decodeBegin(const uint8_t *src, uint32_t n, uint32_t pos0)   // EntroCoders.hx:229-233
    {
        data = src; len = n; overrun = false; wbase = 0x80000000u;
        reinit(pos0);
        nDec = 0;
    }
    __device__ __forceinline__ int get() { if (overrun) fail = true; return (int)(x & 4095u); }   // decGet, ANS.hx:35
    // decAdvance, ANS.hx:37-44.  Everything a valid stream does between two symbols, behind ONE test: the state is normalised
    // (2^23 <= x < 2^31), the interval is sane, and the two bytes a renormalisation can need (x' >= freq * 2^11 >= 2^11, so two
    // bytes reach 2^23) are in the window and in the stream.  Then `while (x < L) x = x << 8 | byte` is two compares and two
    // selects -- no loop and no per-byte bounds checks on the symbol chain.
    __device__ __forceinline__ bool try_advance(int start, int freq)
    {
        const uint32_t off = pos - wbase, f = x & 4095u;
        if ((x - ANS_L) < (0x80000000u - ANS_L) && off < ANS_WIN - 2u && pos + 2u <= len &&
            (uint32_t)(freq - 1) < (uint32_t)ANS_SCALE && f >= (uint32_t)start && !overrun) {
            const uint32_t wa = sa_wk + (uint32_t)offsetof(AnsWork, win) + off;
            const uint32_t b0 = a_lds8(wa), b1 = a_lds8(wa + 1u);
            const uint32_t v = (uint32_t)freq * (x >> 12) + (f - (uint32_t)start);
            const uint32_t v1 = (v << 8) | b0, v2 = (v << 16) | (b0 << 8) | b1;
            const bool one = v < ANS_L, two = v < (ANS_L >> 8);
            x = two ? v2 : (one ? v1 : v);
            pos += (one ? 1u : 0u) + (two ? 1u : 0u);
            return true;
        }
        return false;
    }
    __device__ __forceinline__ void advance(int start, int freq) { if (!try_advance(start, freq)) advance_slow(start, freq); }
    __device__ __forceinline__ void advance_slow(int start, int freq)
    {
        if (pos - wbase >= ANS_WIN - 2u && pos + 2u <= len) {          // only the window was in the way: move it and try again
            refill(pos);
            if (try_advance(start, freq)) return;
        }
        const int32_t r = (int32_t)x;
        if (r >= 0 && (uint32_t)freq <= (uint32_t)ANS_SCALE && (r & 4095) >= start) {
            // every state a valid stream reaches: the product fits 32 bits (freq <= 2^12, r >> 12 < 2^19)
            uint32_t v = (uint32_t)freq * ((uint32_t)r >> 12) + (uint32_t)((r & 4095) - start);
            int guard = 0;
            while (v < ANS_L) {
                if (overrun || ++guard > 8) { fail = true; break; }
                v = (v << 8) | rbyte();
            }
            x = v;
            return;
        }
        long long v = (long long)freq * (r >> 12) + (r & 4095) - start;
        int guard = 0;
        while (v < (long long)ANS_L) {
            if (overrun || ++guard > 8) { fail = true; break; }
            v = (int32_t)(((uint32_t)v << 8) | rbyte());
        }
        x = (uint32_t)v;
    }
    __device__ __forceinline__ void count()                           // EntroCoders.hx:249-253
    {
        nDec++; nsym++;
        if (nDec == ANS_B) { reinit(pos); nDec = 0; }
    }

    template <int N>
    __device__ __forceinline__ void fx_renew(FxTab<N> &t)                             // FixedSizeRansCtx.renew, ANS.hx:128-144
    {
        const int lane = (int)lane_id();
        const int fr = ANS_SCALE / N, c0 = fr - (fr >> 1);
        for (int i = lane; i < FxTab<N>::NP + 8; i += 32) {
            t.cum[i] = i < N ? (uint16_t)(i * fr) : (uint16_t)0xFFFFu;  // sentinels behind the table end every forward scan
            if (i < FxTab<N>::NP) { t.fr[i] = i < N ? (uint16_t)fr : 0; t.cnt[i] = i < N ? (uint16_t)c0 : 0; }
        }
        const int i = (lane * 128) / fr;                              // the symbol whose interval holds 128 * lane
        if (i < N) t.dec[lane] = (uint8_t)i;
        if constexpr (FxTab<N>::NLUT != 0)
            for (int b = lane; b < FxTab<N>::NLUT; b += 32) t.lut[b] = (typename FxTab<N>::lut_t)((16 * b) / fr);
        if (lane == 0) t.cntsum = (uint32_t)(c0 * N);
    }
    // EntroCoders.hx:216-227.  `ponly` = where the tables only P frames use live right now: the shared-memory copy, or (I-frame
    // kernel, which maps only the I-frame prefix) the stream's state in HBM.
    __device__ __forceinline__ void renewI(AnsSmall *ponly)
    {
        gen = gen + 1;
        AnsSmall &s = *small;
        for (int i = 0; i < 6; i++) { fx_renew(s.ntab[i]); fx_renew(s.ptypetab[i]); }
        AnsSmall &p = *ponly;
        fx_renew(p.xxtab); fx_renew(p.ntab2); fx_renew(p.bttab);
        for (int i = 0; i < 4; i++) fx_renew(p.sxytab[i]);
        fx_renew(p.mvtab[0]); fx_renew(p.mvtab[1]);
        __syncwarp();
    }
    __device__ __forceinline__ void renewI() { renewI(small); }

    // FixedSizeRansCtx.decode + incrCnt on a shared-memory table (decodeF, EntroCoders.hx:271-280).
    // The symbol chain of a warp is latency-bound: every warp collective costs 25-45 cycles on it, every shared-memory round
    // trip ~30, every branch ~15 (tools/microbench).  So the common case is ONE branch and no collectives: all loads the symbol
    // can need are issued before the test (table sums, the cumulative frequencies or the lut entry, the next two bitstream
    // bytes), the test checks everything the straight-line code assumes (no rebuild due, state normalised, bytes in the
    // window, no reload of the state due), and every lane then runs the same scalar search -- small tables compare all
    // cumulative frequencies at once and pick freq / count out of registers, big ones start from lut[f >> 4] and scan forward
    // four entries at a time (sentinels end the scan).  Anything else -- the periodic rebuild (ANS.hx:89-102, every ~128
    // symbols of a table), the end of the stream, corrupt states -- takes decodeF_slow.
    static __device__ __forceinline__ uint32_t pick16(const uint4 &v, int c)      // u16 element c (0..7) of a packed vector
    {
        const uint32_t lo = (c & 2) ? v.y : v.x, hi = (c & 2) ? v.w : v.z;
        const uint32_t w = (c & 4) ? hi : lo;
        return (c & 1) ? (w >> 16) : (w & 0xFFFFu);
    }
    static __device__ __forceinline__ int rank8(const uint4 &v, uint32_t f)       // how many of elements 1..7 are <= f
    {
        // a balanced tree: written as one sum the compiler emits a chain of seven dependent conditional increments
        const int a = (int)((v.x >> 16) <= f) + (int)((v.y & 0xFFFFu) <= f), b = (int)((v.y >> 16) <= f) + (int)((v.z & 0xFFFFu) <= f);
        const int c = (int)((v.z >> 16) <= f) + (int)((v.w & 0xFFFFu) <= f), d = (int)((v.w >> 16) <= f);
        return (a + b) + (c + d);
    }
    template <int N>
    __device__ __forceinline__ int decodeF(uint32_t small_off)
    {
        typedef FxTab<N> T;
        const uint32_t ta = sa_small + small_off;                      // the table as a shared-window address
        T &t = *reinterpret_cast<T *>(reinterpret_cast<uint8_t *>(small) + small_off);
        const uint32_t off = pos - wbase, offc = min(off, ANS_WIN - 2u);
        const uint32_t wa = sa_wk + (uint32_t)offsetof(AnsWork, win) + offc;
        const uint32_t b0 = a_lds8(wa), b1 = a_lds8(wa + 1u);
        const uint32_t cs = a_lds32(ta + (uint32_t)offsetof(T, cntsum));
        const uint32_t f = x & 4095u;
        uint4 vc, vf, vn;
        uint32_t c0 = 0;
        if constexpr (N <= 8) {
            vc = a_lds128(ta + (uint32_t)offsetof(T, cum));            // cum[0..7]; entries >= N are 0xFFFF
            vf = a_lds128(ta + (uint32_t)offsetof(T, fr));
            vn = a_lds128(ta + (uint32_t)offsetof(T, cnt));
        } else if constexpr (T::NLUT != 0) {
            // the symbol that holds slot f & ~15: at most 15 short
            c0 = sizeof(typename T::lut_t) == 1 ? a_lds8(ta + (uint32_t)offsetof(T, lut) + (f >> 4))
                                                 : a_lds16(ta + (uint32_t)offsetof(T, lut) + 2u * (f >> 4));
        }
        const bool ok = (x - ANS_L) < (0x80000000u - ANS_L) && off < ANS_WIN - 2u && pos + 2u <= len && !overrun &&
                        cs + 32u <= (uint32_t)ANS_SCALE && nDec + 1 != ANS_B;
        if (!ok) return decodeF_slow(t);
        int c;
        uint32_t freq, cumf, cn;
        if constexpr (N <= 8) {
            c = rank8(vc, f);
            freq = pick16(vf, c); cumf = pick16(vc, c); cn = pick16(vn, c);
        } else {
            c = (int)c0;
            uint32_t cp = ta + (uint32_t)offsetof(T, cum) + 2u * (uint32_t)(c + 1);
            for (;;) {
                const uint32_t a0 = a_lds16(cp), a1 = a_lds16(cp + 2u), a2 = a_lds16(cp + 4u), a3 = a_lds16(cp + 6u);
                const int k = ((int)(a0 <= f) + (int)(a1 <= f)) + ((int)(a2 <= f) + (int)(a3 <= f));
                c += k;
                if (k < 4) break;
                cp += 8u;
            }
            freq = a_lds16(ta + (uint32_t)offsetof(T, fr) + 2u * (uint32_t)c);
            cumf = a_lds16(ta + (uint32_t)offsetof(T, cum) + 2u * (uint32_t)c);
            cn = a_lds16(ta + (uint32_t)offsetof(T, cnt) + 2u * (uint32_t)c);
        }
        __syncwarp();                                                  // every lane has read the table
        a_sts16(ta + (uint32_t)offsetof(T, cnt) + 2u * (uint32_t)c, (cn + 16u) & 0xFFFFu);   // every lane stores the same values
        a_sts32(ta + (uint32_t)offsetof(T, cntsum), cs + 16u);
        const uint32_t v = freq * (x >> 12) + (f - cumf);              // decAdvance, ANS.hx:37-44 (see try_advance)
        const uint32_t v1 = (v << 8) | b0, v2 = (v << 16) | (b0 << 8) | b1;
        const bool one = v < ANS_L, two = v < (ANS_L >> 8);
        x = two ? v2 : (one ? v1 : v);
        pos += (one ? 1u : 0u) + (two ? 1u : 0u);
        nDec++; nsym++;
        return c;
    }
    // the same symbol without assumptions: rebuild when due, generic advance (window refills, failures), state reloads
    template <int N>
    __device__ __forceinline__ int decodeF_slow(FxTab<N> &t)
    {
        const uint32_t cntsum = t.cntsum + 16u;
        if (cntsum + 16u > (uint32_t)ANS_SCALE) return decodeF_rebuild(t);
        const uint32_t f = (uint32_t)get();
        int c = 0;
        if constexpr (FxTab<N>::NLUT != 0) c = (int)t.lut[f >> 4];
        const uint16_t *cp = t.cum + c + 1;
        for (;;) {
            const uint32_t a0 = cp[0], a1 = cp[1], a2 = cp[2], a3 = cp[3];
            const int k = (int)(a0 <= f) + (int)(a1 <= f) + (int)(a2 <= f) + (int)(a3 <= f);
            c += k;
            if (k < 4) break;
            cp += 4;
        }
        const int freq = t.fr[c], cumf = t.cum[c];
        const uint32_t cn = t.cnt[c];
        __syncwarp();
        t.cnt[c] = (uint16_t)(cn + 16u);
        t.cntsum = cntsum;
        __syncwarp();
        advance(cumf, freq);
        count();
        return c;
    }
    // the symbol that triggers the rebuild (fx_rebuild_symbol: out of line, shared by every call site of a table size)
    template <int N>
    __device__ __forceinline__ int decodeF_rebuild(FxTab<N> &t)
    {
        const uint2 r = fx_rebuild_symbol<N>(&t, get());
        advance((int)(r.y >> 16), (int)(r.y & 0xFFFFu));
        count();
        return (int)r.x;
    }

    __device__ __forceinline__ void slot_writeback(int slot, int tag)
    {
        const int lane = (int)lane_id();
        const uint4 *s4 = reinterpret_cast<const uint4 *>(&wk->cache[slot]);
        uint4 *gh = hdrs + (size_t)tag * (ANS_HDR_BYTES / 16);
        uint4 *gb = bodies + (size_t)tag * (ANS_BODY_BYTES / 16);
        if (lane < 4) gh[lane] = s4[lane];
        gb[lane] = s4[4 + lane];
    }
    __device__ __forceinline__ void flush_slots()
    {
        for (int k = 0; k < ANS_CACHE_SLOTS; k++) {
            const int t = __shfl_sync(FULLMASK, my_tag, k);
            if (t >= 0) slot_writeback(k, t);
        }
        my_tag = -1; my_age = 0;
        __syncwarp();
    }

    // Cx7 = FixedSizeRansCtx(256) in global memory: the register path of the fixed tables.  hv / bv = the header words
    // (lanes 0-3) and the first 512 body bytes (cumFreq, 8 per lane) already loaded by the caller.
    __device__ __forceinline__ int decode_cx7(uint4 *gh, uint4 *gb, const uint4 &hv, const uint4 &bv, int f)
    {
        const int lane = (int)lane_id();
        uint4 *sh4 = reinterpret_cast<uint4 *>(wk->big);               // the header's working copy (decTable lives in it)
        if (lane < 4) sh4[lane] = hv;
        __syncwarp();
        CxHdr &H = *reinterpret_cast<CxHdr *>(wk->big);
        uint32_t cum[8], fr[8], cnt[8];
        {
            const uint4 fv = gb[32 + lane], cv = gb[64 + lane];
            auto unpack = [&](const uint4 &w, uint32_t (&o)[8]) {
                o[0] = w.x & 0xFFFFu; o[1] = w.x >> 16; o[2] = w.y & 0xFFFFu; o[3] = w.y >> 16;
                o[4] = w.z & 0xFFFFu; o[5] = w.z >> 16; o[6] = w.w & 0xFFFFu; o[7] = w.w >> 16;
            };
            unpack(bv, cum); unpack(fv, fr); unpack(cv, cnt);
        }
        uint32_t cntsum = H.cntsum;
        int freq, cumf, owner; bool rebuilt;
        const int c = fx_core<256, 8>(cum, fr, cnt, H.dec, cntsum, f, freq, cumf, rebuilt, owner);
        auto pack = [&](const uint32_t (&v)[8]) {
            return make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
        };
        if (rebuilt) { gb[lane] = pack(cum); gb[32 + lane] = pack(fr); gb[64 + lane] = pack(cnt); }
        else if (lane == owner) gb[64 + lane] = pack(cnt);
        if (lane == 0) H.cntsum = cntsum;
        __syncwarp();
        if (lane < (rebuilt ? 4 : 1)) gh[lane] = sh4[lane];
        __syncwarp();
        advance(cumf, freq);
        return c;
    }

    __device__ __forceinline__ int decodeClr(int cxi)                                 // EntroCoders.hx:235-255
    {
        const int lane = (int)lane_id();
#ifdef JSP_PROFILE_SECTIONS
        long long _at = clock64();
#endif
        const int f = get();
        // ---- find the context in the shared-memory cache, or bring it in ----
        const uint32_t hit = __ballot_sync(FULLMASK, lane < ANS_CACHE_SLOTS && my_tag == cxi);
        tick++;
        int slot;
        if (hit) {
            slot = __ffs(hit) - 1;
        } else {
            uint4 *gh = hdrs + (size_t)cxi * (ANS_HDR_BYTES / 16);
            uint4 *gb = bodies + (size_t)cxi * (ANS_BODY_BYTES / 16);
            const uint4 bv = gb[lane];
            uint4 hv = make_uint4(0, 0, 0, 0);
            if (lane < 4) hv = gh[lane];
            const uint32_t h_gen = __shfl_sync(FULLMASK, hv.x, 0), h_w1 = __shfl_sync(FULLMASK, hv.y, 0);
            if (h_gen == gen && ((h_w1 >> 16) & 0xFFu) == CXK_7) {     // big contexts are decoded where they live
                const int c7 = decode_cx7(gh, gb, hv, bv, f);
                JSP_AT(4)
                count();
                return c7;
            }
            const uint32_t key = lane < ANS_CACHE_SLOTS ? ((my_age << 4) | (uint32_t)lane) : 0xFFFFFFFFu;
            slot = (int)(__reduce_min_sync(FULLMASK, key) & 15u);
            const int old = __shfl_sync(FULLMASK, my_tag, slot);
            if (old >= 0) slot_writeback(slot, old);
            __syncwarp();
            uint4 *s4 = reinterpret_cast<uint4 *>(&wk->cache[slot]);
            if (lane < 4) s4[lane] = hv;
            s4[4 + lane] = bv;
            if (lane == slot) my_tag = cxi;
            __syncwarp();
        }
        if (lane == slot) my_age = tick;
        if (hit) { JSP_AT(0) } else { JSP_AT(5) }
        AnsSlot &S = wk->cache[slot];
        CxHdr &H = S.hdr;
        // everything the Cx4 fast path reads, loaded at once BEFORE the kind is tested (one shared-memory round trip on the
        // symbol chain instead of two), plus the next two bitstream bytes for its renormalisation
        const uint32_t ss = sa_wk + (uint32_t)offsetof(AnsWork, cache) + (uint32_t)slot * (uint32_t)sizeof(AnsSlot);   // the slot, shared-window address
        const uint2 hw = a_lds64(ss);                                   // gen | d, kind, maxpos
        const uint32_t symw = a_lds32(ss + (uint32_t)offsetof(AnsSlot, body));
        const uint2 fw = a_lds64(ss + (uint32_t)offsetof(AnsSlot, body) + (uint32_t)B_SC_FR);
        const uint32_t woff = pos - wbase, woffc = min(woff, ANS_WIN - 2u);
        const uint32_t wb0 = a_lds8(sa_wk + (uint32_t)offsetof(AnsWork, win) + woffc), wb1 = a_lds8(sa_wk + (uint32_t)offsetof(AnsWork, win) + woffc + 1u);
        const int kind = hw.x == gen ? (int)((hw.y >> 16) & 0xFFu) : CXK_NONE;
        int c;
        // (Measured dead end: keeping the interval arithmetic of this path as a "decode-ready" record beside the context -- four
        // scaled starts and frequencies, rebuilt after every update -- shortens the symbol chain to a load and four compares, but
        // the rebuild costs as many instructions as it saves and an in-order warp pays for them all the same: 574 -> 627 cycles
        // per symbol.)
        // ---- fast path: a Cx4 context (<= 4 symbols met: the common case on screen content) that HITS one of its
        //      symbols.  SmallContext.decodeSC (ANS.hx:263-309) for S = 4, run by every lane from three shared-memory
        //      words -- no divergence; lane 0 stores the two words that change. ----
        if (kind == CXK_4) {
            const uint32_t f01 = fw.x, f23 = fw.y;
            const int d = (int)(hw.y & 0xFFFFu), mp = (int)(hw.y >> 24);
            const int q0 = (int)(f01 & 0xFFFFu), q1 = (int)(f01 >> 16), q2 = (int)(f23 & 0xFFFFu), q3 = (int)(f23 >> 16);
            const int tot0 = q0 + q1 + q2 + q3 + 256 - d;                                      // Cx4.decode, :320
            // `while (tot <= PROB_SCALE / 2) { tot <<= 1; shift++; }` in closed form (tot0 is 257 .. 4096 + 3 * 50)
            const int hb = 31 - __clz(tot0);
            const int shift = max(0, 11 - hb + ((tot0 & (tot0 - 1)) == 0 ? 1 : 0));
            const int tot = tot0 << shift;
            const int sf = f >> shift;
            const int bonus = (ANS_SCALE - tot) >> shift;
            // symbols (unused slots: 256, which nothing reaches) and this call's frequencies (the bonus goes to maxpos)
            const int s0 = (int)(symw & 0xFFu), s1 = d > 1 ? (int)((symw >> 8) & 0xFFu) : 256,
                      s2 = d > 2 ? (int)((symw >> 16) & 0xFFu) : 256, s3 = d > 3 ? (int)(symw >> 24) : 256;
            const int r0 = (q0 + (mp == 0 ? bonus : 0)) & 0xFFFF, r1 = (q1 + (mp == 1 ? bonus : 0)) & 0xFFFF,
                      r2 = (q2 + (mp == 2 ? bonus : 0)) & 0xFFFF, r3 = (q3 + (mp == 3 ? bonus : 0)) & 0xFFFF;
            // interval starts: every symbol below s_i that has not been met owns one unit (decodeSC, :274-299)
            // (a1 = e0 + s1 - s0 - 1 and so on, written as prefix sums so that the four starts do not form a chain of eight adds)
            const int r01 = r0 + r1;
            const int a0 = s0, a1 = r0 + s1 - 1, a2 = r01 + s2 - 2, a3 = r01 + r2 + s3 - 3;
            const int e0 = a0 + r0, e1 = a1 + r1, e2 = a2 + r2, e3 = a3 + r3;
            const bool h0 = sf >= a0 && sf < e0, h1 = d > 1 && sf >= a1 && sf < e1,
                       h2 = d > 2 && sf >= a2 && sf < e2, h3 = d > 3 && sf >= a3 && sf < e3;
            if (h0 || h1 || h2 || h3) {
                const int hitpos = h0 ? 0 : (h1 ? 1 : (h2 ? 2 : 3));
                c = h0 ? s0 : (h1 ? s1 : (h2 ? s2 : s3));
                const int hstart = h0 ? a0 : (h1 ? a1 : (h2 ? a2 : a3));
                const int hfr = h0 ? r0 : (h1 ? r1 : (h2 ? r2 : r3));
                int n0 = (q0 + (h0 ? 50 : 0)) & 0xFFFF, n1 = (q1 + (h1 ? 50 : 0)) & 0xFFFF,
                    n2 = (q2 + (h2 ? 50 : 0)) & 0xFFFF, n3 = (q3 + (h3 ? 50 : 0)) & 0xFFFF;
                const int fh = h0 ? n0 : (h1 ? n1 : (h2 ? n2 : n3));
                const int fm = mp == 0 ? n0 : (mp == 1 ? n1 : (mp == 2 ? n2 : n3));
                const int nmp = (hitpos != mp && fh > fm) ? hitpos : mp;                      // :291-292
                if (tot0 + 50 + 50 > ANS_SCALE) {                                             // rescale, :254-261
                    n0 = d > 0 ? (n0 - (n0 >> 1)) & 0xFFFF : n0; n1 = d > 1 ? (n1 - (n1 >> 1)) & 0xFFFF : n1;
                    n2 = d > 2 ? (n2 - (n2 >> 1)) & 0xFFFF : n2; n3 = d > 3 ? (n3 - (n3 >> 1)) & 0xFFFF : n3;
                }
                __syncwarp();                                          // every lane has read the slot
                // every lane stores the same three values (no divergent region on the chain)
                a_sts32(ss + (uint32_t)offsetof(AnsSlot, body) + (uint32_t)B_SC_FR, (uint32_t)n0 | ((uint32_t)n1 << 16));
                a_sts32(ss + (uint32_t)offsetof(AnsSlot, body) + (uint32_t)B_SC_FR + 4u, (uint32_t)n2 | ((uint32_t)n3 << 16));
                a_sts8(ss + (uint32_t)offsetof(CxHdr, maxpos), (uint32_t)nmp);
                __syncwarp();
                // decAdvance + the symbol count (ANS.hx:37-44, EntroCoders.hx:249-253): straight-line when the state is normalised,
                // the two bytes are in the window and no state reload is due -- see try_advance; else the generic pair
                const uint32_t start = (uint32_t)(hstart << shift), freq = (uint32_t)(hfr << shift), f12 = x & 4095u;
                if ((x - ANS_L) < (0x80000000u - ANS_L) && woff < ANS_WIN - 2u && pos + 2u <= len && !overrun &&
                    (freq - 1u) < (uint32_t)ANS_SCALE && f12 >= start && nDec + 1 != ANS_B) {
                    const uint32_t v = freq * (x >> 12) + (f12 - start);
                    const uint32_t v1 = (v << 8) | wb0, v2 = (v << 16) | (wb0 << 8) | wb1;
                    const bool one = v < ANS_L, two = v < (ANS_L >> 8);
                    x = two ? v2 : (one ? v1 : v);
                    pos += (one ? 1u : 0u) + (two ? 1u : 0u);
                    nDec++; nsym++;
                    JSP_AT(1)
                    return c;
                }
                advance((int)start, (int)freq);
                JSP_AT(1)
                count();
                return c;
            }
        }
        // ---- generic path: lane 0 does the bookkeeping of the small kinds in place on the cached slot ----
        int wb;
        if (kind >= CXK_4) {
            if (lane == 0) {
                CxRes r; r.c = 0; r.freq = 1; r.cum = 0;
                wk->res_wb = cx::decode_small(H, S.body, wk->big, f, r);
                wk->res_c = r.c; wk->res_freq = r.freq; wk->res_cum = r.cum;
            }
            __syncwarp();
            c = wk->res_c; wb = wk->res_wb;
            const int freq = wk->res_freq, cumf = wk->res_cum;
            advance(cumf, freq);
            if (c > 255) { fail = true; c &= 255; }                   // escape interval past symbol 255: not a valid stream
            JSP_AT(2)
        } else {
            c = (int)rbyte();                                         // Rans.raw, ANS.hx:46-48
            if (lane == 0) wk->res_wb = cx::update_raw(H, S.body, wk->big, c, f0, gen);
            __syncwarp();
            wb = wk->res_wb;
            JSP_AT(3)
        }
        if (wb < 0) {
            // the context has just become a Cx7 (built in `big`): it moves out to global memory and leaves the cache
            uint4 *gh = hdrs + (size_t)cxi * (ANS_HDR_BYTES / 16);
            uint4 *gb = bodies + (size_t)cxi * (ANS_BODY_BYTES / 16);
            const uint4 *big4 = reinterpret_cast<const uint4 *>(wk->big);
            const uint4 *s4 = reinterpret_cast<const uint4 *>(&S);
            gb[lane] = big4[lane]; gb[32 + lane] = big4[32 + lane]; gb[64 + lane] = big4[64 + lane];
            if (lane < 4) gh[lane] = s4[lane];
            if (lane == slot) { my_tag = -1; my_age = 0; }
            __syncwarp();
        }
        count();
        return c;
    }

    // ---- per-frame set-up / tear-down: the first `small_bytes` of the stream's small tables move to shared memory and back ----
    __device__ __forceinline__ void begin_iframe(const SpJob &J)
    {
        renewI(small_bytes == (uint32_t)sizeof(AnsSmall) ? small : &reinterpret_cast<AnsState *>(J.state)->small);
    }
    __device__ __forceinline__ void open(const SpJob &J, AnsSmall *small_sh, AnsWork *work_sh, uint32_t nbytes)
    {
        AnsState *st = reinterpret_cast<AnsState *>(J.state);
        small = small_sh; wk = work_sh; small_bytes = nbytes;
        sa_small = a_opaque_smem(small_sh); sa_wk = a_opaque_smem(work_sh);
        hdrs = st->hdrs; bodies = st->bodies; gen = st->gen;
        f0 = (J.flags & SPJ_ANS_V3) ? 64 : 32;                       // Cx6.f0, EntroCoders.hx:210 / ScreenPressor.hx:69-72
        fail = false; overrun = false; x = 0; data = J.src; len = J.len; pos = 0; wbase = 0x80000000u; nDec = 0; nsym = 0;
        my_tag = -1; my_age = 0; tick = 0;
#ifdef JSP_PROFILE_SECTIONS
        for (int k = 0; k < 16; k++) aprof[k] = 0;
#endif
        const uint4 *g = reinterpret_cast<const uint4 *>(&st->small);
        uint4 *s = reinterpret_cast<uint4 *>(small_sh);
        for (int i = (int)lane_id(); i < (int)(nbytes / 16); i += 32) s[i] = g[i];
        __syncwarp();
    }
    // the interface the kernels of sp2_decode.cu use for both coders (small / wk / small_bytes are set by the kernel)
    __device__ __forceinline__ void open_iframe(const SpJob &J) { open(J, small, wk, small_bytes); }
    __device__ __forceinline__ void open_pframe(const SpJob &J) { open(J, small, wk, small_bytes); }
    __device__ __forceinline__ void close_frame(const SpJob &J) { close(J); }
    __device__ __forceinline__ void close(const SpJob &J)
    {
        AnsState *st = reinterpret_cast<AnsState *>(J.state);
        flush_slots();
        __syncwarp();
        uint4 *g = reinterpret_cast<uint4 *>(&st->small);
        const uint4 *s = reinterpret_cast<const uint4 *>(small);
        for (int i = (int)lane_id(); i < (int)(small_bytes / 16); i += 32) g[i] = s[i];
        if (lane_id() == 0) st->gen = gen;
#ifdef JSP_PROFILE_SECTIONS
        if (lane_id() == 0) for (int k = 0; k < 16; k++) atomicAdd(&g_ans_prof[k], (unsigned long long)aprof[k]);
#endif
    }

    __device__ __forceinline__ bool decodeBool()                                      // EntroCoders.hx:259-269
    {
        const int f = get();
        const bool flag = f >= (ANS_SCALE >> 1);
        advance(flag ? ANS_SCALE >> 1 : 0, ANS_SCALE >> 1);
        count();
        return flag;
    }
    __device__ __forceinline__ int decodeN(int ptype) { return decodeF<256>((uint32_t)offsetof(AnsSmall, ntab) + (uint32_t)ptype * (uint32_t)sizeof(FxTab<256>)); }
    __device__ __forceinline__ int decodeP(int ptype) { return decodeF<6>((uint32_t)offsetof(AnsSmall, ptypetab) + (uint32_t)ptype * (uint32_t)sizeof(FxTab<6>)); }
    __device__ __forceinline__ int decodeX() { return decodeF<256>((uint32_t)offsetof(AnsSmall, xxtab)); }
    __device__ __forceinline__ int decodeBT() { return decodeF<5>((uint32_t)offsetof(AnsSmall, bttab)); }
    __device__ __forceinline__ int decodeBN() { return decodeF<256>((uint32_t)offsetof(AnsSmall, ntab2)); }
    __device__ __forceinline__ int decodeSXY(int n) { return decodeF<16>((uint32_t)offsetof(AnsSmall, sxytab) + (uint32_t)n * (uint32_t)sizeof(FxTab<16>)); }
    __device__ __forceinline__ int decodeMX() { return decodeF<512>((uint32_t)offsetof(AnsSmall, mvtab)); }
    __device__ __forceinline__ int decodeMY() { return decodeF<512>((uint32_t)offsetof(AnsSmall, mvtab) + (uint32_t)sizeof(FxTab<512>)); }
};

// one frame of one rANS stream; `sm` = this warp's shared memory (first-generation kernel: one warp does everything)
__device__ __forceinline__ void sp_ans_run(const SpJob &J, AnsShared &sm, uint32_t *ring, uint32_t *ptile)
{
    const int lane = (int)lane_id();
    AnsCoder ec;
    ec.open(J, &sm.small, &sm.work, (uint32_t)sizeof(AnsSmall));
    uint32_t bits = 0;
    if (J.flags & SPJ_RENEW) {
        ec.renewI();
    } else if (J.flags & SPJ_IFRAME) {
        sp_decode_iframe(ec, J, ring);
        bits |= ST_CHANGED;
    } else {
        sp_decode_pframe(ec, J, bits, ptile);
    }
    ec.close(J);
    if (ec.failed()) {
        bits = ST_ERROR;
        if (!(J.flags & SPJ_RENEW)) sp_undo_frame(J, (J.flags & SPJ_IFRAME) != 0);
    }
    __syncwarp();
    if (lane == 0) { if (bits) atomicOr(J.status, bits); if (J.symbols) *J.symbols = ec.nsym; }
    sp_signal_done(J);
}

}  // namespace jsp
