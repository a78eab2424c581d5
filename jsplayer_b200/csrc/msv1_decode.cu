// msv1_decode.cu -- Microsoft Video 1 (CRAM/MSVC) batch decode for sm_100a.
//
// Replaces the serial block walk of reference src/MSVideo1.hx:106-209 (RGB555) and :293-393
// (8-bit palettised) with ONE kernel that reads every compressed byte once and writes every
// output pixel once (HBM-bound; no tensor-core work exists on this path).
//
// The bitstream of every frame of the launch is cut into 4 KiB tiles; one CTA owns one tile:
//   scan phase  (per tile, in shared memory)
//     1. every lane walks its 16-word segment BACKWARDS and gets, for each possible entry offset
//        0..8 (an opcode is at most 9 words), the offset at which the opcode chain leaves the segment
//        -> a 9-nibble map in one 64-bit register.  Opcode boundaries depend on bytes only.
//     2. maps are chained lane -> group of 8 lanes -> warp -> tile with warp shuffles: lane 8g + e follows entry e through
//        the 8 segments of group g and records the entry into each (maps are rarely constant -- never when every opcode
//        has 3 words -- so nothing waits for a chain to self-synchronise).
//     3. tiles of one frame are chained with a single-pass decoupled look-back (one 64-bit state word
//        per tile); tiles are ticketed tile-major over all frames so predecessors are long finished.
//     4. lanes walk forward from their now-known entry, count blocks (skip runs weigh n), and a
//        warp-shuffle prefix sum + a second look-back give every opcode its absolute 4x4-block index;
//        a per-opcode (position, block) table lands in shared memory.
//   fill phase: one thread per coded block, consecutive threads = consecutive blocks, four 16-byte
//     streaming stores per block (a warp writes 4 x 512 contiguous bytes); skip runs are copied from
//     the previous picture warp-per-run with 16-byte loads/stores.
//
// JavaScript semantics of the reference on truncated input are reproduced exactly (see
// oracle/msvideo1_oracle.c): bytes past the end read as `undefined`.
#include "common.cuh"

namespace jsp {
namespace {

typedef unsigned long long u64;

constexpr u64 FLAG_AGG  = 1ull << 62;
constexpr u64 FLAG_INCL = 2ull << 62;
constexpr u64 FLAG_MASK = 3ull << 62;
constexpr u64 MAP_TERM  = 0xFull << 60;           // nibble 15 maps to 15: "chain terminated" is absorbing
constexpr u64 MAP_IDENT = 0x876543210ull | MAP_TERM;
constexpr u64 ONES9     = 0x111111111ull;
constexpr uint32_t TERM = 15;
// Watchdog of the two look-back spins.  Forward progress is by construction (tickets hand tiles out in table order, so a
// tile's predecessors are held by CTAs that are resident or finished and only ever wait for lower tickets), but a spin with no
// bound would turn any future scheduling mistake into a silent hang of the stream.  A predecessor that has not published after
// this many polls (seconds; a tile takes microseconds) fails the FRAME (ST_ERROR, nothing more is decoded into it) instead.
// -DJSP_SPIN_LIMIT=n overrides (tests/test_msv1_gpu.py builds nothing special: the limit is always on).
#ifndef JSP_SPIN_LIMIT
#define JSP_SPIN_LIMIT (1u << 26)
#endif
constexpr uint32_t MSV1_SPIN_LIMIT = JSP_SPIN_LIMIT;

__device__ __forceinline__ uint32_t nib(u64 m, uint32_t e) { return (uint32_t)(m >> (4 * e)) & 15u; }

template <int NENT>
__device__ __forceinline__ bool map_const(u64 m, uint32_t &c)
{
    constexpr u64 MASK = (1ull << (4 * NENT)) - 1;
    c = (uint32_t)m & 15u;
    return (m & MASK) == ((ONES9 & MASK) * c);
}

// (A then B)(e) = B[A[e]]
template <int NENT>
__device__ __forceinline__ u64 compose(u64 A, u64 B)
{
    uint32_t c;
    if (map_const<NENT>(B, c)) return B;
    u64 r = MAP_TERM;
#pragma unroll
    for (int e = 0; e < NENT; e++) r |= (u64)nib(B, nib(A, e)) << (4 * e);
    return r;
}

__device__ __forceinline__ u64 ld_state(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 shfl64(u64 v, int src)
{
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}

// MSVideo1.hx:211-214 fromRGB15, for a colour in the low / high half of a 32-bit word (no extraction needed)
// The three fields are masked on the ALU pipe and positioned with integer multiply-adds on the otherwise idle
// FMA pipe (the kernel is ALU-pipe bound, profiles/r01_msv1_ncu.md); the fields never overlap, so + is |.
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t mad_hi_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// Five instructions per colour: red and blue are expanded by ONE multiply -- (R << 10 | B) * 0x208 = R << 19 | B << 3 plus two
// stray copies (B << 9, R << 13) that lie between the wanted fields and cannot carry into them (their sum stays below 2^19),
// so one mask cleans up -- and green is positioned by the multiply-add that also merges it.  The colour in the HIGH half of a
// word takes the same five through multiply-high (no shift down first).
__device__ __forceinline__ uint32_t rgb15(uint32_t c)
{
    const uint32_t rb = ((c & 0x7C1Fu) * 0x208u) & 0x00F800F8u;
    return mad_u32(c & 0x3E0u, 64u, rb);
}
__device__ __forceinline__ uint32_t rgb15_hi(uint32_t x)
{
    const uint32_t rb = __umulhi(x & 0x7C1F0000u, 0x02080000u) & 0x00F800F8u;
    return mad_hi_u32(x & 0x03E00000u, 1u << 22, rb);
}

__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b, uint32_t cap)
{
    uint32_t s = a + b;
    return s < cap ? s : cap;
}

// explicit global-space streaming accesses (the pointers come out of a descriptor, so the compiler would
// otherwise emit generic-space ST/LD)
__device__ __forceinline__ void st_global_cs(void *p, const uint4 &v)
{
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_global_cs(const void *p)
{
    uint4 v;
    asm volatile("ld.global.cs.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(__cvta_generic_to_global(p)));
    return v;
}

struct Smem {
    alignas(16) uint8_t stage[2][MSV1_STAGE_BYTES];   // double-buffered bitstream tiles (cp.async prefetch)
    alignas(16) Msv1Frame fd[2];                        // frame descriptors of the two staged tiles
    uint32_t tk[4];                                     // tickets handed from thread 0 to the CTA
    uint32_t blk[MSV1_TILE_WORDS];     // absolute block index of opcode k
    uint16_t pos[MSV1_TILE_WORDS];     // word position in the tile | 0x8000 for copy runs
    uint16_t runs[MSV1_TILE_WORDS];    // opcode indices of the short skip runs
    int32_t  pal[256];
    u64      wmap[4];
    uint32_t wentry[4];
    uint32_t wops[4];
    uint32_t wblk[4];
    uint32_t ticket;
    uint32_t first_block;
    uint32_t nruns;
    uint32_t big_blk0;                 // first block of a "rest of frame" copy, or 0xFFFFFFFF
    uint32_t flags;
    uint32_t terminated;
    uint32_t any_runs;
};

// 16 reconstructed pixels of one coded block from 8 quadrant colours (2-colour and 1-colour blocks are
// replicated into the same form).  bit i of `flags` = 1 selects the odd colour of the pair.
// (Measured alternative: the kernel's ALU pipe is 73 % busy at 68 % issue-active while the FMA pipe idles, so the select was
// rewritten as FMA-pipe arithmetic -- bit i to bit 31 by a multiply, down by a multiply-high, pixel = even + bit * (odd - even)
// by a multiply-add: three FMA instructions instead of a bit test and a select.  2.20 -> 2.27 ms on C2: the extra issue slot
// per pixel costs more than the ALU relief gains -- both limits are close.)
// DISP (fused display store, JSP_BATCH_DISPLAY): the colours arrive already converted; with `flip` picture row y is
// stored in row Y-1-y (the caller's render-time flip, Main.hx:946).
template <bool DISP>
__device__ __forceinline__ void store_block(int32_t *out, uint32_t X, uint32_t by, uint32_t bx,
                                            const uint32_t (&col)[8], uint32_t flags, bool vec_ok,
                                            uint32_t Y = 0, bool flip = false)
{
    // one global-space byte address, advanced by the row pitch (the compiler otherwise keeps a 64-bit pixel index AND rebuilds
    // the pointer from it for every row), and ONE test of vec_ok per block
    uint32_t first = by * 4u * X + bx * 4u;           // pictures stay below 2^32 pixels
    long long pitch = (long long)X * 4;
    if constexpr (DISP) { if (flip) { first = (Y - 1u - by * 4u) * X + bx * 4u; pitch = -pitch; } }
    unsigned long long a = (unsigned long long)__cvta_generic_to_global(out) + (unsigned long long)first * 4u;
    uint32_t px[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const int q = ((r & 2) << 1) + (x & 2);
            px[r][x] = ((flags >> (4 * r + x)) & 1u) ? col[q + 1] : col[q];
        }
    }
    if (vec_ok) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(px[r][0]), "r"(px[r][1]), "r"(px[r][2]), "r"(px[r][3]) : "memory");
            a += (unsigned long long)pitch;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            asm volatile("st.global.b32 [%0], %1;" ::"l"(a), "r"(px[r][0]) : "memory");
            asm volatile("st.global.b32 [%0+4], %1;" ::"l"(a), "r"(px[r][1]) : "memory");
            asm volatile("st.global.b32 [%0+8], %1;" ::"l"(a), "r"(px[r][2]) : "memory");
            asm volatile("st.global.b32 [%0+12], %1;" ::"l"(a), "r"(px[r][3]) : "memory");
            a += (unsigned long long)pitch;
        }
    }
}

// (the four rows of a block are copied in buffer order, so a flipped block simply starts three rows further up)
template <bool DISP>
__device__ __forceinline__ void copy_block(int32_t *out, const int32_t *prev, uint32_t X, uint32_t by,
                                           uint32_t bx, bool vec_ok, uint32_t Y = 0, bool flip = false)
{
    size_t off = (size_t)by * 4u * X + bx * 4u;
    if constexpr (DISP) { if (flip) off = (size_t)(Y - 4u - by * 4u) * X + bx * 4u; }
    constexpr uint32_t ZERO = DISP ? 0xFF000000u : 0u;             // a pixel nobody wrote, as the canvas shows it
    int32_t *d = out + off;
    if (vec_ok) {
        uint4 v[4];
        if (prev) {
            const int32_t *s = prev + off;
#pragma unroll
            for (int r = 0; r < 4; r++) v[r] = ld_global_cs(s + (size_t)r * X);
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++) v[r] = make_uint4(ZERO, ZERO, ZERO, ZERO);
        }
#pragma unroll
        for (int r = 0; r < 4; r++) st_global_cs(d + (size_t)r * X, v[r]);
    } else {
        for (int r = 0; r < 4; r++)
            for (int x = 0; x < 4; x++) d[(size_t)r * X + x] = prev ? prev[off + (size_t)r * X + x] : (int32_t)ZERO;
    }
}

// Manager.fill_bitmap_data (Manager.hx:363-381): 0x00RRGGBB -> the Int32 view of canvas bytes R,G,B,A (alpha 255)
__device__ __forceinline__ uint32_t disp_px(uint32_t c) { return __byte_perm(c, 0xFFu, 0x4012); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(__cvta_generic_to_global(gsrc)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Starts the copy of one bitstream tile (+32 B look-ahead, zero-filled past the end of the frame) and of its frame
// descriptor into shared memory.  16-byte aligned full chunks travel as cp.async (no registers, no waiting);
// the rare unaligned / trailing chunks are assembled from 32-bit loads.
__device__ __forceinline__ void stage_tile(const Msv1Tile &e, const Msv1Frame *frames, uint8_t *dst, Msv1Frame *fd, uint32_t tid)
{
    const uint8_t *g = e.src;
    const int avail = (int)e.avail;
    const bool aligned = (reinterpret_cast<uintptr_t>(g) & 15u) == 0;
    uint4 *s4 = reinterpret_cast<uint4 *>(dst);
    for (int c = tid; c < MSV1_STAGE_BYTES / 16; c += MSV1_THREADS) {
        const int b0 = c * 16;
        if (aligned && b0 + 16 <= avail) { cp_async16(s4 + c, g + b0); continue; }
        uint4 v = make_uint4(0, 0, 0, 0);
        if (b0 < avail) {
            // unaligned source or the chunk straddles the end: assemble from aligned 32-bit words
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int o = b0 + 4 * k;
                uint32_t val = 0;
                if (o < avail) {
                    const uintptr_t addr = reinterpret_cast<uintptr_t>(g + o);
                    const uint32_t *ga = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
                    const uint32_t sh = (uint32_t)(addr & 3u) * 8u;
                    const uint32_t lo = __ldg(ga);
                    const uint32_t hi = (sh != 0 && o + 4 - (int)(sh >> 3) < avail) ? __ldg(ga + 1) : 0u;
                    val = __funnelshift_r(lo, hi, sh);
                    const int rem = avail - o;
                    if (rem < 4) val &= (1u << (8 * rem)) - 1u;
                }
                w[k] = val;
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        s4[c] = v;
    }
    static_assert(sizeof(Msv1Frame) % 16 == 0, "Msv1Frame is copied in 16-byte units");
    if (tid < sizeof(Msv1Frame) / 16)
        cp_async16(reinterpret_cast<uint4 *>(fd) + tid, reinterpret_cast<const uint4 *>(frames + e.frame) + tid);
}

// Persistent CTAs: a CTA works through tickets (tiles in table order).  While it decodes tile T0 the bitstream and
// descriptor of T1 are in flight (cp.async into the other shared-memory buffer), the table entry of T2 is in
// flight in registers and the atomic for T3's ticket has been issued -- the four dependent global latencies a tile
// needs before its first instruction are all hidden behind the previous tile's work.  `depth0` (small launches:
// one tile per CTA, nothing held back) keeps the look-back chain of a single frame short.
template <bool IS8, bool DISP>
__global__ void __launch_bounds__(MSV1_THREADS, 8)
msv1_decode_kernel(const Msv1Frame *__restrict__ frames, const Msv1Tile *__restrict__ tiles, uint32_t n_tiles, int depth0,
                   u64 *__restrict__ tile_map, u64 *__restrict__ tile_cnt, unsigned int *__restrict__ ticket)
{
    constexpr int NENT = IS8 ? 5 : 9;      // possible entry offsets into a segment (longest opcode: 5 / 9 words)
    __shared__ Smem sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t FULL = 0xffffffffu;
    const uint32_t NONE = 0xFFFFFFFFu;

    // ---- work assignment: tickets hand tiles out in table order, so every tile a CTA may wait for is held by a
    //      resident CTA that only ever waits for lower tickets (forward progress) ----
    if (tid == 0) {
        sm.tk[0] = atomicAdd(ticket, 1u);
        sm.tk[1] = depth0 ? NONE : atomicAdd(ticket, 1u);
        sm.tk[2] = depth0 ? NONE : atomicAdd(ticket, 1u);
    }
    __syncthreads();
    uint32_t tk0 = sm.tk[0], tk1 = sm.tk[1], tk2 = sm.tk[2];
    if (tk0 >= n_tiles) return;
    Msv1Tile e0 = tiles[tk0], e1 = e0, e2 = e0;
    if (tk1 < n_tiles) e1 = tiles[tk1];
    if (tk2 < n_tiles) e2 = tiles[tk2];
    stage_tile(e0, frames, sm.stage[0], &sm.fd[0], tid);
    cp_async_commit();
    int cur = 0;
  for (;;) {
    uint32_t tk3 = NONE;
    if (tid == 0) {
        if (!depth0 && tk2 < n_tiles) tk3 = atomicAdd(ticket, 1u);       // consumed at the end of this tile
        sm.nruns = 0; sm.big_blk0 = 0xFFFFFFFFu; sm.flags = 0; sm.terminated = 0; sm.any_runs = 0;
    }
    if (tk1 < n_tiles) stage_tile(e1, frames, sm.stage[cur ^ 1], &sm.fd[cur ^ 1], tid);
    cp_async_commit();
    cp_async_wait<1>();                    // everything but the group just committed has landed: this tile is in
    __syncthreads();
    uint8_t *const sbytes = sm.stage[cur];
    const Msv1Frame &F = sm.fd[cur];
    const uint32_t len = F.len, X = F.X, nbx = F.nbx, nblocks = F.nblocks;
    const uint32_t Y = F.Y;
    const bool flip = DISP && (F.flags & MSV1_F_FLIP) != 0;
    constexpr uint32_t ZERO = DISP ? 0xFF000000u : 0u;
    const uint32_t tile = (uint32_t)((e0.src - F.src) / MSV1_TILE_BYTES);
    const uint32_t tile_byte0 = tile * MSV1_TILE_BYTES;
    const uint32_t n_words = (len + 1u) >> 1;                      // an odd trailing byte is a half word (see below)
    const uint32_t tile_words = n_words > tile * MSV1_TILE_WORDS
                                    ? min((uint32_t)MSV1_TILE_WORDS, n_words - tile * MSV1_TILE_WORDS) : 0u;
    const bool vec_ok = ((X & 3u) == 0) && ((reinterpret_cast<uintptr_t>(F.out) & 15u) == 0) &&
                        ((reinterpret_cast<uintptr_t>(F.prev) & 15u) == 0);
    // The two look-backs below usually find their predecessor's state already published (tiles are ticketed tile-major, the
    // predecessor started a whole round of frames earlier): load both words NOW, so that the L2 round trips run under the scan
    // instead of in the two single-thread sections the other 127 threads wait for.  A word that is not there yet is polled later.
    u64 pf_map = 0, pf_cnt = 0;
    if (tid == 0 && tile > 0) {
        pf_map = ld_state(tile_map + F.state_base + tile - 1);
        pf_cnt = ld_state(tile_cnt + F.state_base + tile - 1);
    }
    if (IS8) {
        for (int i = tid; i < 256; i += MSV1_THREADS) {
            const uint32_t c = F.pal ? (uint32_t)F.pal[i] : 0u;
            sm.pal[i] = (int32_t)(DISP ? disp_px(c) : c);
        }
        __syncthreads();
    }
    // An odd frame length leaves a half word {a, undefined}.  The reference then takes the 1-colour
    // branch with (undefined<<8)+a (MSVideo1.hx:171-173) / pal[a] (:353); patching the missing high byte
    // to 0x80 selects exactly that class with exactly that colour (bit 15 is ignored by fromRGB15).
    if ((len & 1u) && len >= tile_byte0 && len - tile_byte0 < (uint32_t)MSV1_STAGE_BYTES) {
        if (tid == 0) sbytes[len - tile_byte0] = 0x80;
        __syncthreads();
    }

    // ---- scan 1: backward walk of the lane's 16-word segment.  R[p] describes the opcode chain that starts at
    //      word p: bits 0-15 = the words where its opcodes start, bits 16-19 = where it leaves the segment
    //      (offset into the next one, or 15 = terminated).  All indexing is static after unrolling. ----
    uint32_t R[MSV1_SEG_WORDS];
    uint32_t skipmask = 0, termmask = 0;      // per word: "is a copy run" / "ends the frame" if an opcode starts there
    u64 M;
    {
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(sbytes) + tid * 8;
        const uint4 v0 = *reinterpret_cast<const uint4 *>(sw);
        const uint4 v1 = *reinterpret_cast<const uint4 *>(sw + 4);
        const uint32_t r[9] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, sw[8]};
#define JSP_NEXT(L) ((p + (L) < MSV1_SEG_WORDS) ? R[(p + (L)) & (MSV1_SEG_WORDS - 1)] : (uint32_t)((p + (L) - MSV1_SEG_WORDS) << 16))
        if constexpr (!IS8) {
            // RGB555: an opcode's length depends on two bits -- bit 15 of its own word (1 colour or skip run: 1 word) and bit 15 of
            // the NEXT word (8 colours: 9 words, else 3) -- and a skip run on its high byte.  Both are classified for all 16 words
            // at once: PRMT gathers the four high bytes of two registers, a multiply gathers one bit per byte into a nibble.
            uint32_t G = 0, S = 0;                 // bit p: word p has bit 15 set / word p is a skip run ((b & 0xFC) == 0x84, :131)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t hb = __byte_perm(r[2 * i], r[2 * i + 1], 0x7531);
                G |= ((((hb >> 7) & 0x01010101u) * 0x01020408u) >> 24) << (4 * i);
                const uint32_t z = (hb ^ 0x84848484u) & 0xFCFCFCFCu;             // a zero byte = a skip-run word
                const uint32_t nz = (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & 0x80808080u;
                S |= (((((nz ^ 0x80808080u) >> 7)) * 0x01020408u) >> 24) << (4 * i);
            }
            G |= ((r[8] >> 15) & 1u) << 16;
            uint32_t T = 0;                        // skip count 0: the rest of the frame is copied (:132, :124) -- word 0x8400
            for (uint32_t m = S; m; m &= m - 1) {
                const int p = __ffs(m) - 1;
                if (reinterpret_cast<const uint16_t *>(sbytes)[tid * MSV1_SEG_WORDS + p] == 0x8400u) T |= 1u << p;
            }
            skipmask = S; termmask = T;
            if (T == 0) {
#pragma unroll
                for (int p = MSV1_SEG_WORDS - 1; p >= 0; --p) {
                    const bool one = (G >> p) & 1u, eight = (G >> (p + 1)) & 1u;
                    const uint32_t nxt = one ? JSP_NEXT(1) : (eight ? JSP_NEXT(9) : JSP_NEXT(3));
                    R[p] = nxt | (1u << p);
                }
            } else {
#pragma unroll
                for (int p = MSV1_SEG_WORDS - 1; p >= 0; --p) {
                    const bool one = (G >> p) & 1u, eight = (G >> (p + 1)) & 1u;
                    uint32_t nxt = one ? JSP_NEXT(1) : (eight ? JSP_NEXT(9) : JSP_NEXT(3));
                    if ((T >> p) & 1u) nxt = TERM << 16;
                    R[p] = nxt | (1u << p);
                }
            }
        } else {
#pragma unroll
        for (int p = MSV1_SEG_WORDS - 1; p >= 0; --p) {
            uint32_t a, b;
            if (p & 1) { a = (r[p >> 1] >> 16) & 0xFFu; b = r[p >> 1] >> 24; }
            else       { a = r[p >> 1] & 0xFFu; b = (r[p >> 1] >> 8) & 0xFFu; }
            bool isrun = (b & 0xFCu) == 0x84u;                     // skip run (MSVideo1.hx:315)
            bool term = (b == 0x84u) && (a == 0u);                 // skip count 0: rest of the frame is copied (:124)
            const bool t8 = (a + b == 0u);                         // terminator (MSVideo1.hx:313)
            term = term || t8; isrun = isrun || t8;
            uint32_t nxt = (b < 0x80u) ? JSP_NEXT(2) : (b >= 0x90u ? JSP_NEXT(5) : JSP_NEXT(1));
            if (term) nxt = TERM << 16;
            R[p] = nxt | (1u << p);
            if (isrun) skipmask |= 1u << p;
            if (term) termmask |= 1u << p;
        }
        }
#undef JSP_NEXT
        uint32_t lo = 0;
#pragma unroll
        for (int e = 0; e < 8; e++) lo |= (R[e] >> 16) << (4 * e);
        M = ((u64)(R[8] >> 16) << 32) | lo | MAP_TERM;
    }

    // ---- scan 2: chain the maps.  A segment's map is rarely constant (14 % of the segments of the quoted mix, none where every
    //      opcode has 3 words), so entries are not handed from lane to lane (one lane per shuffle round: 12 + 6 rounds per tile on
    //      the mix, 31 on all-2-colour input) but tracked: the warp's 32 segments form 4 groups of 8; lane 8g + e follows entry e
    //      (and, every lane of the group, entry 8) through the 8 segments of group g in 8 steps, keeping the entry INTO each
    //      segment as a nibble.  Group exits chain to the warp's map here and, once the warp's own entry is known (look-back
    //      below), to each group's entry; a lane's entry is then one nibble of the record of the lane that followed that value. ----
    const uint32_t grp = lane & 24u, sub = lane & 7u;
    uint32_t t_e = sub, t_8 = 8u, inter = 0, inter8 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u64 Mi = shfl64(M, (int)grp + i);
        inter |= t_e << (4 * i); inter8 |= t_8 << (4 * i);
        t_e = nib(Mi, t_e); t_8 = nib(Mi, t_8);            // TERM (15) is absorbing: every map carries MAP_TERM
    }
    {
        uint32_t cur = lane < 9u ? lane : 0u;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const uint32_t a = __shfl_sync(FULL, t_e, 8 * g + (int)(cur & 7u));
            const uint32_t b = __shfl_sync(FULL, t_8, 8 * g);
            cur = cur == TERM ? TERM : (cur < 8u ? a : b);
        }
        const uint32_t lo = __reduce_or_sync(FULL, lane < (uint32_t)(NENT < 8 ? NENT : 8) ? cur << (4 * lane) : 0u);
        const uint32_t hi = __shfl_sync(FULL, cur, 8);
        if (lane == 0) sm.wmap[warp] = ((u64)(NENT > 8 ? hi : 0u) << 32) | lo | MAP_TERM;
    }
    __syncthreads();
    // the tile's map = the four warp maps chained: entry e of lanes 0..NENT-1 of warp 0 walks through them, the nibbles are
    // gathered with two OR-reductions (one thread composing 64-bit maps nibble by nibble cost 100+ instructions that the
    // other 127 threads waited for)
    u64 A = 0;
    if (warp == 0) {
        uint32_t e = lane < (uint32_t)NENT ? lane : 0u;
#pragma unroll
        for (int w = 0; w < 4; w++) e = nib(sm.wmap[w], e);
        const uint32_t lo = __reduce_or_sync(FULL, lane < (uint32_t)(NENT < 8 ? NENT : 8) ? e << (4 * lane) : 0u);
        const uint32_t hi = __shfl_sync(FULL, e, 8);
        A = ((u64)(NENT > 8 ? hi : 0u) << 32) | lo | MAP_TERM;
    }
    if (tid == 0) {
        u64 *slot = tile_map + F.state_base + tile;
        uint32_t c; const bool aconst = map_const<NENT>(A, c);
        if (tile + 1 < F.n_tiles) st_state(slot, aconst ? (FLAG_INCL | c) : (FLAG_AGG | (A & 0xFFFFFFFFFull)));
        // look back for this tile's entry offset
        uint32_t e0 = 0;
        bool stalled = false;
        if (tile > 0) {
            u64 acc = MAP_IDENT;
            int j = (int)tile - 1;
            for (;;) {
                u64 st = (j == (int)tile - 1) ? pf_map : 0ull;
                uint32_t spins = 0;
                while ((st & FLAG_MASK) == 0 && ++spins < MSV1_SPIN_LIMIT) st = ld_state(tile_map + F.state_base + j);
                if ((st & FLAG_MASK) == 0) { stalled = true; e0 = TERM; break; }      // watchdog: see MSV1_SPIN_LIMIT
                if ((st & FLAG_MASK) == FLAG_INCL) { e0 = nib(acc, (uint32_t)st & 15u); break; }
                acc = compose<NENT>((st & 0xFFFFFFFFFull) | MAP_TERM, acc);
                uint32_t cc;
                if (map_const<NENT>(acc, cc)) { e0 = cc; break; }
                if (j == 0) { e0 = nib(acc, 0); break; }
                --j;
            }
        }
        if (!aconst && tile + 1 < F.n_tiles) st_state(slot, FLAG_INCL | nib(A, e0));
        uint32_t e = e0;
        for (int w = 0; w < 4; w++) { sm.wentry[w] = e; e = nib(sm.wmap[w], e); }
        if (e0 == TERM) sm.terminated = 1;
        if (stalled) atomicOr(F.status, ST_ERROR);
    }
    __syncthreads();
    uint32_t entry;
    {
        uint32_t eg = sm.wentry[warp], mine = eg;          // eg: the entry into group g (the same in every lane)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            if ((uint32_t)g == (lane >> 3)) mine = eg;
            const uint32_t a = __shfl_sync(FULL, t_e, 8 * g + (int)(eg & 7u));
            const uint32_t b = __shfl_sync(FULL, t_8, 8 * g);
            eg = eg == TERM ? TERM : (eg < 8u ? a : b);
        }
        const uint32_t pk = __shfl_sync(FULL, inter, (int)(grp + (mine & 7u)));
        entry = mine == TERM ? TERM : ((mine < 8u ? pk : inter8) >> (4 * sub)) & 15u;
    }

    // ---- scan 3: the chain of the true entry gives this lane's opcode starts; count its blocks ----
    const int seg_lim = (int)min((uint32_t)MSV1_SEG_WORDS,
                                 tile_words > tid * MSV1_SEG_WORDS ? tile_words - tid * MSV1_SEG_WORDS : 0u);
    uint32_t starts = 0;
#pragma unroll
    for (int e = 0; e < NENT; e++) if (entry == (uint32_t)e) starts = R[e];
    starts &= (1u << seg_lim) - 1u;                                // words past the end of the frame hold no opcodes
    const uint32_t runs = starts & skipmask;
    const bool lterm = (starts & termmask) != 0;
    const uint32_t nops = __popc(starts);
    uint32_t blocks = nops - __popc(runs);
    if (runs) {
        uint32_t m = runs & ~termmask;
        while (m) {
            const int p = __ffs(m) - 1; m &= m - 1;
            const uint8_t *o = sbytes + (tid * MSV1_SEG_WORDS + p) * 2;
            blocks += (((uint32_t)o[1] - 0x84u) << 8) | o[0];
        }
        if (lterm) blocks = nblocks;
        blocks = min(blocks, nblocks);
    }
    // ---- scan 4: block / opcode prefix sums (warp shuffles + look-back) ----
    uint32_t op_incl = nops, blk_incl = blocks;
    const bool anyrun = __any_sync(FULL, runs != 0);
    if (!anyrun) {                                                 // no copy run in this warp's 32 segments: one block per opcode
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o2 = __shfl_up_sync(FULL, op_incl, d);
            if (lane >= (uint32_t)d) op_incl += o2;
        }
        blk_incl = min(op_incl, nblocks);                          // (a chain of saturating adds = the saturated sum)
    } else {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o2 = __shfl_up_sync(FULL, op_incl, d);
            const uint32_t b2 = __shfl_up_sync(FULL, blk_incl, d);
            if (lane >= (uint32_t)d) { op_incl += o2; blk_incl = sat_add(blk_incl, b2, nblocks); }
        }
    }
    if (lane == 31) { sm.wops[warp] = op_incl; sm.wblk[warp] = blk_incl; }
    {
        const bool anyterm = __any_sync(FULL, lterm);
        if (lane == 0) { if (anyterm) sm.terminated = 1; if (anyrun) sm.any_runs = 1; }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < 4; w++) tot = sat_add(tot, sm.wblk[w], nblocks);
        u64 *slot = tile_cnt + F.state_base + tile;
        const bool publish = tile + 1 < F.n_tiles;
        uint32_t first = 0;
        if (tile > 0) {
            if (publish) st_state(slot, FLAG_AGG | tot);
            int j = (int)tile - 1;
            for (;;) {
                // the prefetched word may hold the predecessor's AGGREGATE where its inclusive sum has arrived since: both are valid
                u64 st = (j == (int)tile - 1) ? pf_cnt : 0ull;
                uint32_t spins = 0;
                while ((st & FLAG_MASK) == 0 && ++spins < MSV1_SPIN_LIMIT) st = ld_state(tile_cnt + F.state_base + j);
                if ((st & FLAG_MASK) == 0) { atomicOr(F.status, ST_ERROR); first = nblocks; break; }   // watchdog: decode nothing
                first = sat_add(first, (uint32_t)st, nblocks);
                if ((st & FLAG_MASK) == FLAG_INCL || j == 0) break;
                --j;
            }
        }
        if (publish) st_state(slot, FLAG_INCL | sat_add(first, tot, nblocks));
        sm.first_block = first;
    }
    __syncthreads();
    const uint32_t first_block = sm.first_block;
    const bool any_runs = sm.any_runs != 0;
    uint32_t tot_ops = 0, tot_blk = 0, op_base = op_incl - nops, blk_base;
    {
        uint32_t wb = 0;
        for (uint32_t w = 0; w < 4; w++) {
            if (w == warp) { op_base += tot_ops; wb = tot_blk; }
            tot_ops += sm.wops[w]; tot_blk = sat_add(tot_blk, sm.wblk[w], nblocks);
        }
        const uint32_t prev_incl = __shfl_up_sync(FULL, blk_incl, 1);
        blk_base = sat_add(wb, lane ? prev_incl : 0u, nblocks);
    }
    const uint32_t incl_block = sat_add(first_block, tot_blk, nblocks);
    const bool last_tile = tile + 1 == F.n_tiles;

    uint32_t myflags = 0;
    // (the streaming loads / stores are asm statements that clobber memory: a descriptor field read inside a loop over blocks
    //  would be re-read from shared memory after every block)
    int32_t *const out_pic = F.out;
    const int32_t *const prev_pic = F.prev;
    const uint32_t insign_blocks = F.insign_blocks;
    if (first_block < nblocks) {
        // ---- per-opcode table: word position (and, only when the tile has copy runs, the block index;
        //      otherwise opcode k of the tile is simply block first_block + k) ----
        if (!any_runs) {
            uint32_t m = starts, opi = op_base;
            while (m) {
                const int p = __ffs(m) - 1; m &= m - 1;
                sm.pos[opi++] = (uint16_t)(tid * MSV1_SEG_WORDS + p);
            }
        } else {
            uint32_t m = starts, opi = op_base, blk = sat_add(first_block, blk_base, nblocks);
            while (m) {
                const int p = __ffs(m) - 1; m &= m - 1;
                const bool run = (skipmask >> p) & 1u, big = (termmask >> p) & 1u;
                sm.pos[opi] = (uint16_t)((tid * MSV1_SEG_WORDS + p) | (run ? 0x8000u : 0u));
                sm.blk[opi] = blk;
                // only the chain's FIRST terminator counts.  With the grouped entry tracking TERM is absorbing, so no lane behind
                // a terminator parses the trailing bytes any more (the lane-to-lane hand-over did: an entry fixed by a constant map
                // before the terminator was known); atomicMin stays as the order-independent way to publish it
                if (big) { atomicMin(&sm.big_blk0, blk); blk = nblocks; }
                else if (run) {
                    const uint8_t *o = sbytes + (tid * MSV1_SEG_WORDS + p) * 2;
                    sm.runs[atomicAdd(&sm.nruns, 1u)] = (uint16_t)opi;
                    blk = sat_add(blk, (((uint32_t)o[1] - 0x84u) << 8) | o[0], nblocks);
                } else blk = sat_add(blk, 1u, nblocks);
                opi++;
            }
        }
        __syncthreads();

        // ---- fill: one thread per coded block, branch-free over the three block classes ----
        const uint32_t inv_nbx = F.inv_nbx;
        for (uint32_t op = tid; op < tot_ops; op += MSV1_THREADS) {
            const uint32_t pw = sm.pos[op];
            if (pw & 0x8000u) continue;
            const uint32_t blk = any_runs ? sm.blk[op] : first_block + op;
            if (blk >= nblocks) continue;                          // opcodes past the last block are never read
            uint32_t by = __umulhi(blk, inv_nbx), bx = blk - by * nbx;
            if (bx >= nbx) { by++; bx -= nbx; }
            const uint32_t abs0 = tile_byte0 + pw * 2;             // byte offset of the opcode in the frame
            uint32_t col[8], flags;
            if (abs0 + (IS8 ? 10u : 18u) <= len) {
                // the opcode as aligned 32-bit words, shifted down by 16 bits when it starts on an odd word
                const uint32_t *q = reinterpret_cast<const uint32_t *>(sbytes) + (pw >> 1);
                const uint32_t sh = (pw & 1u) * 16u;
                if (IS8) {
                    const uint32_t v0 = __funnelshift_r(q[0], q[1], sh), v1 = __funnelshift_r(q[1], q[2], sh),
                                   v2 = __funnelshift_r(q[2], q[3], sh);
                    const uint32_t a = v0 & 0xFFu, b = (v0 >> 8) & 0xFFu;
                    const bool two = b < 0x80u, eight = b >= 0x90u;
                    flags = two ? (v0 & 0xFFFFu) : (eight ? ((v0 & 0xFFFFu) ^ 0xFFFFu) : 0u);   // MSVideo1.hx:320,337
                    // 2 colours: bit 1 -> pal[first byte], 0 -> pal[second] (:322-323); 1 colour: pal[a] (:353)
                    const uint32_t podd = (uint32_t)sm.pal[two || eight ? ((v0 >> 16) & 0xFFu) : a];
                    const uint32_t peven = (uint32_t)sm.pal[two ? (v0 >> 24) : (eight ? ((v0 >> 16) & 0xFFu) : a)];
                    // (for 8 colours podd/peven are placeholders; the real pairs follow)
#pragma unroll
                    for (int k = 0; k < 8; k += 2) { col[k] = peven; col[k + 1] = podd; }
                    if (eight) {                                   // colours in stream order (:338-340)
                        col[0] = (uint32_t)sm.pal[(v0 >> 16) & 0xFFu]; col[1] = (uint32_t)sm.pal[v0 >> 24];
                        col[2] = (uint32_t)sm.pal[v1 & 0xFFu];         col[3] = (uint32_t)sm.pal[(v1 >> 8) & 0xFFu];
                        col[4] = (uint32_t)sm.pal[(v1 >> 16) & 0xFFu]; col[5] = (uint32_t)sm.pal[v1 >> 24];
                        col[6] = (uint32_t)sm.pal[v2 & 0xFFu];         col[7] = (uint32_t)sm.pal[(v2 >> 8) & 0xFFu];
                    }
                } else {
                    const uint32_t v0 = __funnelshift_r(q[0], q[1], sh), v1 = __funnelshift_r(q[1], q[2], sh),
                                   v2 = __funnelshift_r(q[2], q[3], sh), v3 = __funnelshift_r(q[3], q[4], sh),
                                   v4 = __funnelshift_r(q[4], q[5], sh);
                    const bool one = (v0 & 0x8000u) != 0;          // b >= 0x80: 1 colour = the opcode word (:171-181)
                    const bool eight = !one && (v0 & 0x80000000u); // colour0 bit 15 (:142)
                    flags = one ? 0u : ((v0 & 0xFFFFu) ^ 0xFFFFu); // :136
                    // one expansion for both classes (the compiler turned `one ? rgb15(v0) : rgb15_hi(v0)` into a divergent branch);
                    // a 1-colour block never selects an odd colour (flags = 0), so c1 needs no select
                    const uint32_t c0 = rgb15_hi(one ? v0 << 16 : v0);
                    const uint32_t c1 = rgb15(v1);
                    col[0] = c0; col[1] = c1;
                    col[2] = eight ? rgb15_hi(v1) : c0; col[3] = eight ? rgb15(v2) : c1;
                    col[4] = eight ? rgb15_hi(v2) : c0; col[5] = eight ? rgb15(v3) : c1;
                    col[6] = eight ? rgb15_hi(v3) : c0; col[7] = eight ? rgb15(v4) : c1;
                }
            } else {
                // the opcode runs past the end of the frame: JavaScript `undefined` reads, byte by byte
                const uint8_t *o = sbytes + pw * 2;
                auto rd = [&](uint32_t i) -> int { return abs0 + i < len ? (int)o[i] : -1; };
                auto w16 = [&](uint32_t i) -> uint32_t { return abs0 + i + 1 < len ? (uint32_t)o[i] | ((uint32_t)o[i + 1] << 8) : 0u; };
                const int a = rd(0), b = rd(1);
                flags = 0;
                if (IS8) {
                    auto palu = [&](int ix) -> uint32_t { return ix < 0 ? ZERO : (uint32_t)sm.pal[ix]; };
                    if (b >= 0 && b < 0x80) {
                        flags = ((uint32_t)b << 8) | (uint32_t)a;
                        const uint32_t c1 = palu(rd(2)), c0 = palu(rd(3));
                        for (int k = 0; k < 8; k += 2) { col[k] = c0; col[k + 1] = c1; }
                    } else if (b >= 0x90) {
                        flags = (((uint32_t)b << 8) | (uint32_t)a) ^ 0xFFFFu;
                        for (int k = 0; k < 8; k++) col[k] = palu(rd(2 + k));
                    } else {
                        const uint32_t c = palu(a);
                        for (int k = 0; k < 8; k++) col[k] = c;
                    }
                } else {
                    if (b >= 0 && b < 0x80) {
                        flags = ((((uint32_t)b << 8) | (uint32_t)a) ^ 0xFFFFu);
                        const uint32_t w0 = w16(2);
                        if (w0 & 0x8000u) {
                            for (int k = 0; k < 8; k++) col[k] = rgb15(w16(2 + 2 * k));
                        } else {
                            const uint32_t c0 = rgb15(w0), c1 = rgb15(w16(4));
                            for (int k = 0; k < 8; k += 2) { col[k] = c0; col[k + 1] = c1; }
                        }
                    } else {
                        const uint32_t c = rgb15(b < 0 ? (a < 0 ? 0u : (uint32_t)a) : (((uint32_t)b << 8) | (uint32_t)a));
                        for (int k = 0; k < 8; k++) col[k] = c;
                    }
                }
            }
            if constexpr (DISP && !IS8) {
#pragma unroll
                for (int k = 0; k < 8; k++) col[k] = disp_px(col[k]);
            }
            store_block<DISP>(out_pic, X, by, bx, col, flags, vec_ok, Y, flip);
            myflags |= ST_CHANGED | (by >= insign_blocks ? ST_SIGNIF_ROWS : 0u);
        }

        // ---- skip runs: warp per run, 16-byte copies from the previous picture ----
        const bool precopied = (F.flags & MSV1_F_PRECOPIED) != 0;     // skipped blocks already hold the previous picture
        const uint32_t nruns = precopied ? 0u : sm.nruns;
        for (uint32_t r = warp; r < nruns; r += 4) {
            const uint32_t op = sm.runs[r];
            const uint32_t blk0 = sm.blk[op];
            const uint8_t *o = sbytes + (sm.pos[op] & 0x7FFFu) * 2;
            uint32_t n = (((uint32_t)o[1] - 0x84u) << 8) | o[0];
            n = min(n, nblocks - min(blk0, nblocks));
            for (uint32_t j = lane; j < n; j += 32) {
                const uint32_t blk = blk0 + j, by = blk / nbx, bx = blk - by * nbx;
                copy_block<DISP>(out_pic, prev_pic, X, by, bx, vec_ok, Y, flip);
            }
            if (n && !prev_pic && (F.flags & MSV1_F_HAS_PRED)) myflags |= ST_NEEDS_PREV;
        }
        // ---- "rest of the frame is copied" (skip count 0 / 8-bit terminator): whole CTA ----
        const uint32_t big0 = precopied ? 0xFFFFFFFFu : sm.big_blk0;
        if (big0 < nblocks) {
            for (uint32_t blk = big0 + tid; blk < nblocks; blk += MSV1_THREADS) {
                const uint32_t by = blk / nbx, bx = blk - by * nbx;
                copy_block<DISP>(out_pic, prev_pic, X, by, bx, vec_ok, Y, flip);
            }
            if (!prev_pic && (F.flags & MSV1_F_HAS_PRED)) myflags |= ST_NEEDS_PREV;
        }
    }
    // ---- the bitstream ended before the last block: the reference keeps "reading" undefined bytes, i.e.
    //      1-colour blocks of colour 0 that count as changes (MSVideo1.hx:171-181 with NaN -> 0) ----
    if (last_tile && !sm.terminated && incl_block < nblocks) {
        const uint32_t zero[8] = {ZERO, ZERO, ZERO, ZERO, ZERO, ZERO, ZERO, ZERO};
        for (uint32_t blk = incl_block + tid; blk < nblocks; blk += MSV1_THREADS) {
            const uint32_t by = blk / nbx, bx = blk - by * nbx;
            store_block<DISP>(out_pic, X, by, bx, zero, 0u, vec_ok, Y, flip);
            myflags |= ST_CHANGED | (by >= insign_blocks ? ST_SIGNIF_ROWS : 0u);
        }
    }
    myflags = __reduce_or_sync(FULL, myflags);
    if (lane == 0 && myflags) atomicOr(&sm.flags, myflags);
    uint32_t *const status_ptr = F.status; // read now: the next tile's descriptor is staged over this one right after the barrier
    if (tid == 0) sm.tk[3] = tk3;
    __syncthreads();                       // every thread is done with this tile's shared-memory buffers
    if (tid == 0 && sm.flags) atomicOr(status_ptr, sm.flags);
    // ---- rotate the pipeline ----  (no second barrier: thread 0 rewrites sm.tk[3] only after the next tile's barriers, and
    //      the words it resets at the top of the loop are read by the others only before the barrier above)
    tk0 = tk1; e0 = e1;
    tk1 = tk2; e1 = e2;
    tk2 = sm.tk[3];
    if (tk0 >= n_tiles) break;
    if (tk2 < n_tiles) e2 = tiles[tk2];
    cur ^= 1;
  }
    cp_async_wait<0>();
}

}  // namespace

void launch_msv1_decode(bool is8, bool display, const Msv1Frame *d_frames, const Msv1Tile *d_tiles, uint32_t n_tiles,
                        unsigned long long *d_tile_map, unsigned long long *d_tile_cnt,
                        unsigned int *d_ticket, int sm_count, cudaStream_t st)
{
    if (n_tiles == 0) return;
    // resident CTAs per SM (registers / shared memory allow 8-10): a persistent grid when there is work for several
    // rounds, else one CTA per tile
    const uint32_t capacity = (uint32_t)sm_count * 8u;
    const int depth0 = n_tiles < 6u * capacity ? 1 : 0;
    const uint32_t grid = depth0 ? n_tiles : capacity;
#define JSP_LAUNCH(I8, DISP) msv1_decode_kernel<I8, DISP><<<grid, MSV1_THREADS, 0, st>>>(d_frames, d_tiles, n_tiles, depth0, d_tile_map, d_tile_cnt, d_ticket)
    if (display) { if (is8) JSP_LAUNCH(true, true); else JSP_LAUNCH(false, true); }
    else         { if (is8) JSP_LAUNCH(true, false); else JSP_LAUNCH(false, false); }
#undef JSP_LAUNCH
}

}  // namespace jsp
