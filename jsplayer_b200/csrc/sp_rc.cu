// sp_rc.cu -- ScreenPressor v2 entropy decode on sm_100a: 32-bit byte-wise range decoder with adaptive
// frequency tables.  Replaces reference src/RangeCoder.hx (whole file) and EntroCoderRC
// (src/EntroCoders.hx:31-180) for one stream per warp.
//
//  * coder state (range, code, read position) is replicated in every lane: no broadcasts on the serial chain;
//  * the non-colour tables (ntab 6x257, ptypetab 6x7, xxtab, ntab2, bttab, sxytab 4x17, mvtab 2x513 -- 12.9 KB)
//    live in shared memory for the whole frame and are saved to / restored from the stream's state in HBM;
//  * a colour context is one 1056-byte row in HBM/L2 (256 counts + total + generation tag; the reference's 16
//    group sums are derived data and are not stored).  The warp loads the row with one coalesced 1 KB access,
//    8 counts per lane, and finds the symbol with a shuffle prefix sum + __ballot_sync instead of the reference's
//    linear searches (RangeCoder.hx:58-65, :90-108) -- same symbol, since the search is a pure function of the
//    cumulative counts;
//  * renewI (EntroCoders.hx:81-130) is O(1) for the 12288 colour rows: it bumps a generation number and rows
//    with an older tag read as "all ones" (the reference also resets lazily, :85).
#include "sp_common.cuh"
#include <cstddef>
#include <cstring>

namespace jsp {

constexpr uint32_t RC_TOP = 0x01000000u, RC_BOT = 0x010000u;   // RangeCoder.hx:12-13
constexpr int RC_ROW_STRIDE = 264;                             // u32 per colour row: 256 counts, total, tag, pad
constexpr int RC_ROWS = 3 * 4096;

struct RcSmall {                                               // padded so every table starts 16-byte aligned
    uint32_t ntab[6][260];
    uint32_t xxtab[260];
    uint32_t ntab2[260];
    uint32_t mvtab[2][516];
    uint32_t sxytab[4][20];
    uint32_t ptypetab[6][8];
    uint32_t bttab[8];
};

struct RcState {                                               // per stream, in HBM
    RcSmall small;
    uint32_t gen;                                              // generation of the colour rows (bumped by renewI)
    uint32_t pad[3];
    // followed (separately allocated) by RC_ROWS * RC_ROW_STRIDE u32 of colour rows
    uint32_t *rows;
};

struct RcCoder {
    static constexpr bool kCanDecodeBool = false;              // EntroCoders.hx:178
    RcSmall *sm;                                               // shared memory
    RcState *st;
    uint32_t *rows;
    uint32_t gen;
    uint32_t range, code;
    const uint8_t *data;
    uint32_t len, pos;
    bool poisoned, fail;

    __device__ __forceinline__ bool failed() const { return fail; }

    __device__ __forceinline__ void next_byte()
    {
        if (pos < len) code = (code << 8) | __ldg(data + pos);
        else poisoned = true;                                  // JS: code becomes NaN for good (RangeCoder.hx:41)
        pos++;
    }
    __device__ void decodeBegin(const uint8_t *src, uint32_t n, uint32_t pos0)   // RangeCoder.hx:19-34
    {
        data = src; len = n; code = 0; range = 0xFFFFFFFFu; poisoned = false;
        pos = pos0 + 1;
        next_byte(); next_byte(); next_byte(); next_byte();
    }
    __device__ __forceinline__ uint32_t get_freq(uint32_t tot)                   // RangeCoder.hx:45-49
    {
        if (poisoned) fail = true;
        range = range / tot;
        return poisoned ? 0u : code / range;
    }
    __device__ __forceinline__ void decode(uint32_t cum, uint32_t freq)          // RangeCoder.hx:36-43
    {
        code -= cum * range;
        range *= freq;
        while (range < RC_TOP) { next_byte(); range <<= 8; }
    }

    __device__ void renewI()                                                     // EntroCoders.hx:81-130
    {
        const int lane = (int)lane_id();
        gen = gen + 1;
        for (int t = 0; t < 6; t++) { for (int i = lane; i < 256; i += 32) sm->ntab[t][i] = 1; if (lane == 0) sm->ntab[t][256] = 256; }
        for (int i = lane; i < 256; i += 32) { sm->xxtab[i] = 1; sm->ntab2[i] = 1; }
        for (int t = 0; t < 2; t++) { for (int i = lane; i < 512; i += 32) sm->mvtab[t][i] = 1; if (lane == 0) sm->mvtab[t][512] = 512; }
        if (lane < 16) for (int t = 0; t < 4; t++) sm->sxytab[t][lane] = 1;
        if (lane < 6) for (int t = 0; t < 6; t++) sm->ptypetab[t][lane] = 1;
        if (lane < 5) sm->bttab[lane] = 1;
        if (lane == 0) {
            sm->xxtab[256] = 256; sm->ntab2[256] = 256; sm->bttab[5] = 5;
            for (int t = 0; t < 4; t++) sm->sxytab[t][16] = 16;
            for (int t = 0; t < 6; t++) sm->ptypetab[t][6] = 6;
        }
        __syncwarp();
    }

    // RangeCoder.hx:51-80 for tables of up to 32 symbols: one count per lane
    template <int MAXC>
    __device__ int decode_small(uint32_t *tab, uint32_t step)
    {
        const int lane = (int)lane_id();
        uint32_t tot = tab[MAXC];
        const uint32_t value = get_freq(tot);
        const uint32_t c = lane < MAXC ? tab[lane] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const uint32_t o = __shfl_up_sync(FULLMASK, incl, s); if (lane >= s) incl += o; }
        const int L = __popc(__ballot_sync(FULLMASK, lane < MAXC && value >= incl));
        if (L >= MAXC) { fail = true; return MAXC - 1; }
        const uint32_t cum = __shfl_sync(FULLMASK, incl - c, L), cnt = __shfl_sync(FULLMASK, c, L);
        decode(cum, cnt);
        tot += step;
        uint32_t mine = c + (lane == L ? step : 0u);
        if (tot > RC_BOT) {                                   // :70-77
            mine = lane < MAXC ? (mine >> 1) + 1 : 0u;
            tot = __reduce_add_sync(FULLMASK, mine);
            if (lane < MAXC) tab[lane] = mine;
        } else if (lane == L) tab[lane] = mine;
        if (lane == 0) tab[MAXC] = tot;
        __syncwarp();
        return L;
    }

    // the search + update over K consecutive counts per lane (K = 8: 256 symbols, K = 16: 512 symbols)
    template <int K>
    __device__ __forceinline__ int search_update(uint32_t (&f)[K], uint32_t &tot, uint32_t step, bool &rescaled, int &owner)
    {
        const int lane = (int)lane_id();
        const uint32_t value = get_freq(tot);
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < K; j++) s += f[j];
        uint32_t incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += o; }
        const int L = __popc(__ballot_sync(FULLMASK, value >= incl));
        if (L >= 32) { fail = true; owner = 0; rescaled = false; return 32 * K - 1; }
        // lane L finds the symbol among its K counts
        uint32_t cum = incl - s, cnt = 0; int j = 0;
#pragma unroll
        for (int q = 0; q < K; q++) { if (q == j && value >= cum + f[q] && q < K - 1) { cum += f[q]; j++; } }
#pragma unroll
        for (int q = 0; q < K; q++) if (q == j) cnt = f[q];
        cum = __shfl_sync(FULLMASK, cum, L); cnt = __shfl_sync(FULLMASK, cnt, L); j = __shfl_sync(FULLMASK, j, L);
        decode(cum, cnt);
        if (lane == L) {
#pragma unroll
            for (int q = 0; q < K; q++) if (q == j) f[q] += step;
        }
        tot += step;
        rescaled = tot > RC_BOT;
        if (rescaled) {                                       // RangeCoder.hx:113-127 / :70-77
            uint32_t t = 0;
#pragma unroll
            for (int q = 0; q < K; q++) { f[q] = (f[q] >> 1) + 1; t += f[q]; }
            tot = __reduce_add_sync(FULLMASK, t);
        }
        owner = L;
        return L * K + j;
    }

    // RangeCoder.hx:51-80 for the 256- and 512-symbol tables in shared memory
    template <int K>
    __device__ int decode_big(uint32_t *tab, uint32_t step)
    {
        const int lane = (int)lane_id();
        uint32_t f[K];
        const uint4 *t4 = reinterpret_cast<const uint4 *>(tab) + lane * (K / 4);
#pragma unroll
        for (int q = 0; q < K / 4; q++) { const uint4 v = t4[q]; f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w; }
        uint32_t tot = tab[32 * K];
        bool rescaled; int owner;
        const int c = search_update<K>(f, tot, step, rescaled, owner);
        if (fail) return c;
        if (rescaled || lane == owner) {
            uint4 *o4 = reinterpret_cast<uint4 *>(tab) + lane * (K / 4);
#pragma unroll
            for (int q = 0; q < K / 4; q++) o4[q] = make_uint4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        }
        if (lane == 0) tab[32 * K] = tot;
        __syncwarp();
        return c;
    }

    // RangeCoder.hx:82-130 on a colour row in HBM/L2
    __device__ int decodeClr(int cxi)
    {
        const int lane = (int)lane_id();
        uint32_t *row = rows + (size_t)cxi * RC_ROW_STRIDE;
        uint4 a = reinterpret_cast<const uint4 *>(row)[lane * 2], b = reinterpret_cast<const uint4 *>(row)[lane * 2 + 1];
        const uint2 meta = *reinterpret_cast<const uint2 *>(row + 256);      // total, generation tag
        const bool fresh = meta.y != gen;                                    // not touched since the last renewI
        uint32_t f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t tot = meta.x;
        if (fresh) {
#pragma unroll
            for (int q = 0; q < 8; q++) f[q] = 1;
            tot = 256;
        }
        bool rescaled; int owner;
        const int c = search_update<8>(f, tot, 400u, rescaled, owner);
        if (fail) return c;
        if (fresh || rescaled || lane == owner) {
            reinterpret_cast<uint4 *>(row)[lane * 2] = make_uint4(f[0], f[1], f[2], f[3]);
            reinterpret_cast<uint4 *>(row)[lane * 2 + 1] = make_uint4(f[4], f[5], f[6], f[7]);
        }
        if (lane == 0) *reinterpret_cast<uint2 *>(row + 256) = make_uint2(tot, gen);
        __syncwarp();
        return c;
    }
    __device__ int decodeN(int ptype) { return decode_big<8>(sm->ntab[ptype], 400u); }    // EntroCoders.hx:142-144
    __device__ int decodeP(int ptype) { return decode_small<6>(sm->ptypetab[ptype], 1000u); }
    __device__ int decodeX() { return decode_big<8>(sm->xxtab, 1u); }
    __device__ int decodeBT() { return decode_small<5>(sm->bttab, 10u); }
    __device__ int decodeBN() { return decode_big<8>(sm->ntab2, 20u); }
    __device__ int decodeSXY(int n) { return decode_small<16>(sm->sxytab[n], 100u); }
    __device__ int decodeMX() { return decode_big<16>(sm->mvtab[0], 100u); }
    __device__ int decodeMY() { return decode_big<16>(sm->mvtab[1], 100u); }
    __device__ bool decodeBool() { return false; }
};

namespace {

__global__ void __launch_bounds__(32)
sp_rc_decode_kernel(const SpJob *__restrict__ jobs)
{
    __shared__ RcSmall sm;
    const SpJob J = jobs[blockIdx.x];
    RcState *st = reinterpret_cast<RcState *>(J.state);
    const int lane = (int)lane_id();
    RcCoder ec;
    ec.sm = &sm; ec.st = st; ec.rows = st->rows; ec.gen = st->gen;
    ec.fail = false; ec.poisoned = false; ec.range = 0; ec.code = 0; ec.data = J.src; ec.len = J.len; ec.pos = 0;
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
    if (lane == 0) { st->gen = ec.gen; if (bits) atomicOr(J.status, bits); }
}

}  // namespace

size_t sp_rc_state_bytes() { return (sizeof(RcState) + 255) & ~(size_t)255; }
size_t sp_rc_rows_bytes() { return (size_t)RC_ROWS * RC_ROW_STRIDE * 4; }

// host-side init of one stream's state: generation 1, tables irrelevant until the first I frame
void sp_rc_state_init(void *d_state, void *d_rows, uint32_t gen0, cudaStream_t st)
{
    RcState h;
    memset(&h, 0, sizeof h);
    h.gen = gen0; h.rows = reinterpret_cast<uint32_t *>(d_rows);
    // only the header fields matter; the small tables are rewritten by renewI before use
    cudaStreamSynchronize(st);      // ordered after the memsets queued on st
    cudaMemcpy(reinterpret_cast<char *>(d_state) + offsetof(RcState, gen), &h.gen, sizeof(RcState) - offsetof(RcState, gen),
               cudaMemcpyHostToDevice);
}

void launch_sp_rc(const SpJob *d_jobs, uint32_t n_jobs, cudaStream_t st)
{
    if (n_jobs) sp_rc_decode_kernel<<<n_jobs, 32, 0, st>>>(d_jobs);
}

}  // namespace jsp
