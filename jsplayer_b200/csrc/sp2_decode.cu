// sp2_decode.cu -- second-generation ScreenPressor kernels (range coder; see sp2_rc.cuh for the symbol decoder).
//
//  sp2_rc_i_kernel   coded I frames (ScreenPressor.hx:117-295), one CTA of TWO warps per frame:
//      warp 0 ("E") runs the entropy decode -- the serial chain -- and nothing else: per run it decodes the predictor
//               type, the colour (3 symbols, only for type 0) and the run length, and pushes (type, length, colour) into a
//               shared-memory queue;
//      warp 1 ("R") pops runs and reconstructs pixels: fills, copies from the row above and the gradient predictor, 32
//               pixels per step, written to the picture in HBM and to a shared-memory ring of the last X + 1 pixels.
//      Round 1 did both in one warp, and the run write (~430 cycles: shared-memory round trips, a 5-step shuffle prefix
//      sum for the gradient predictor, two warp barriers) sat in the middle of the symbol chain.  The decoder needs
//      reconstructed pixels only for the colour CONTEXT of the next type-0 run (ScreenPressor.hx:274-275), so that one
//      value is fetched lazily: E waits for R to drain and reads the last pixel from the ring -- by then R has had a whole
//      decodeP of head start.  The gradient predictor needs no prefix sum at all: left + above - aboveleft telescopes
//      along a run to  p[i] = p[start-1] + above[i] - above[start-1]  per byte (R below).
//  sp2_rc_p_kernel   P frames (:302-484) and model resets of flat frames: one warp, round 1's frame loop
//      (sp_common.cuh) on the new symbol decoder.
//
// Separate kernels per (coder, frame type) keep each hot loop small: round 1's single kernel held both coders and both
// frame loops in 68 000 instructions (1.1 MB of SASS) against a 32 KB L1.5 instruction cache.
#include "sp2_rc.cuh"
#include "sp_ans.cuh"
#include <atomic>

namespace jsp {
namespace g2 {

// Optional section timing (JSP_NVCC_EXTRA=-DJSP_SP2_PROF): cycles per warp role, summed over the launch into g_sp2_prof[]
//  0 decodeP  1 decode_rgb  2 decodeN  3 push (incl. back-pressure)  4 drain wait  5 E whole loop  6 runs  7 colour runs
//  8 R wait for an entry  9 R run write  10 R whole loop  11 drains
#ifdef JSP_SP2_PROF
__device__ unsigned long long g_sp2_prof[16];
#define P2_DECL long long _p2[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long _p2t = clock64();
#define P2_T(k) { const long long _n = clock64(); _p2[k] += _n - _p2t; _p2t = _n; }
#define P2_C(k) _p2[k]++;
#define P2_FLUSH if (lane_id() == 0) { for (int _k = 0; _k < 12; _k++) if (_p2[_k]) atomicAdd(&g_sp2_prof[_k], (unsigned long long)_p2[_k]); }
#else
#define P2_DECL
#define P2_T(k)
#define P2_C(k)
#define P2_FLUSH
#endif

// how long the reconstruction warp sleeps between two looks at an empty queue (JSP_SP2_POLL_NS overrides: tuning knob)
__device__ unsigned int g_sp2_poll_ns = 32;

// ---- E -> R run queue (shared memory, single producer / single consumer) -------------------------------------------
// One entry = two words, BOTH carrying generation bits of the slot (run index / RQ_N + 1), so a torn read can never pass
// for a complete entry:  w0 = colour | type << 24 | (gen & 31) << 27,  w1 = length | (gen & 0xFFFF) << 16.
constexpr int RQ_N = 64;
constexpr uint32_t RQ_END = 7;                                   // type 7: end of frame
struct RunQueue {
    alignas(16) uint2 e[RQ_N];
    uint32_t done;                                               // runs R has completed (written by R, read by E)
    uint32_t pad[3];
};

__device__ __forceinline__ void st_volatile_v2(uint2 *p, uint32_t a, uint32_t b)
{
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2 *p)
{
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
// `done` is published with a plain volatile store and read with a plain volatile load.  A release / acquire pair would be
// the textbook choice, but st.release makes the reconstruction warp wait for the acknowledgement of its GLOBAL stores
// (the picture) before every queue step -- hundreds of cycles -- when all that has to be ordered is SHARED memory: ring
// stores before the counter store.  Shared memory is one in-order pipeline per SM, accesses of one warp are performed in
// program order (a __syncwarp() separates the lanes' ring stores from lane 0's counter store), and the reader's ring load
// is issued after its counter load has returned.
__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}

struct Producer {
    RunQueue *q;
    uint32_t issued, seen_done;
    __device__ __forceinline__ void push(uint32_t type, uint32_t n, uint32_t clr)
    {
        while (issued - seen_done >= (uint32_t)RQ_N) seen_done = ld_relaxed(&q->done);     // back-pressure (rare)
        const uint32_t gen = issued / RQ_N + 1u;
        st_volatile_v2(&q->e[issued % RQ_N], (clr & 0xFFFFFFu) | (type << 24) | ((gen & 31u) << 27), n | ((gen & 0xFFFFu) << 16));
        issued++;
    }
    // the same without a branch: the entry is written to the next slot either way (harmless: its generation bits make it valid
    // only once `issued` has moved past it) and `issued` advances only for n > 0
    __device__ __forceinline__ void push_if(bool yes, uint32_t type, uint32_t n, uint32_t clr)
    {
        const uint32_t gen = issued / RQ_N + 1u;
        const uint32_t g0 = yes ? gen : gen - 1u;                                // a skipped entry keeps the slot's OLD generation
        st_volatile_v2(&q->e[issued % RQ_N], (clr & 0xFFFFFFu) | (type << 24) | ((g0 & 31u) << 27), n | ((g0 & 0xFFFFu) << 16));
        issued += yes ? 1u : 0u;
    }
    __device__ __forceinline__ bool full() const { return issued - seen_done >= (uint32_t)RQ_N; }
    __device__ __forceinline__ void wait_room() { while (full()) seen_done = ld_relaxed(&q->done); }
    __device__ __forceinline__ void drain() { while (ld_relaxed(&q->done) != issued) {} seen_done = issued; }
};

// Which of the CTA's two warps decodes?  A warp runs on SM sub-partition (hardware warp id) % 4, a CTA's warps usually get
// ids 2k and 2k + 1, and the entropy warps are the ones that are issue-bound: if it were always the first warp, all of an
// SM's entropy warps would sit on sub-partitions 0 and 2 (measured: 427 cycles per symbol alone, 680 with 7 CTAs per SM,
// 608 with this).  So pair k takes its first warp if (k >> 1) is even and its second otherwise: sub-partitions 0, 2, 1, 3, ...
// Also clears the run queue.  Returns 0 for the entropy warp, 1 for the reconstruction warp -- as a ballot result, i.e. a
// value the compiler can SEE is warp-uniform: a branch on anything derived from threadIdx makes it guard every warp
// collective below with a divergence check (BRA.DIV, ~12 cycles each on the symbol chain).
__device__ __forceinline__ int sp2_pick_roles(RunQueue *rq)
{
    __shared__ uint32_t s_wid[2];
    const int lane = threadIdx.x & 31;
    uint32_t my_wid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(my_wid));
    if (lane == 0) s_wid[threadIdx.x >> 5] = my_wid;
    if (threadIdx.x < RQ_N) rq->e[threadIdx.x] = make_uint2(0u, 0u);
    if (threadIdx.x == 0) rq->done = 0;
    __syncthreads();
    const uint32_t wa = s_wid[0], wb = s_wid[1];
    const bool take_second = ((min(wa, wb) >> 2) & 1u) != 0, am_second = (wa != wb) ? my_wid == max(wa, wb) : threadIdx.x >= 32;
    return __ballot_sync(0xffffffffu, am_second == take_second) ? 0 : 1;
}

// ---- warp 1: pixel reconstruction (ScreenPressor.hx:242-273) -------------------------------------------------------
// The last X + 1 reconstructed pixels: a power-of-two ring in shared memory, or -- pictures too wide for 64 KB, a separate
// instantiation -- the picture itself in HBM (the "ring" index is the pixel index; the same warp wrote those pixels, __syncwarp
// orders them).
// Indices below 0 occur only while the first X + 1 pixels are decoded, where the loaded value is never selected: clamped.
template <bool HBM>
struct RingRef {
    uint32_t *p; uint32_t mask;
    static constexpr bool in_hbm() { return HBM; }
    __device__ __forceinline__ uint32_t ld(long j) const
    {
        if constexpr (HBM) return p[j < 0 ? 0 : j];
        else return p[(uint32_t)j & mask];
    }
    __device__ __forceinline__ void st(long j, uint32_t v) const { if constexpr (!HBM) p[(uint32_t)j & mask] = v; }
};

template <bool HBM>
__device__ __forceinline__ void sp2_recon_iframe(RunQueue *q, const SpJob &J, const RingRef<HBM> ring)
{
    const int lane = (int)lane_id();
    const long X = J.X, end = (long)J.X * J.Y;
    int32_t *dst = J.dst;
    const int chunk = X < 32 ? (int)X : 32;          // a chunk never reads pixels it writes itself (the row above is X away)
    long di = 0;
    uint32_t lastval = 0, consumed = 0;
    const unsigned int poll_ns = g_sp2_poll_ns;
    P2_DECL
#ifdef JSP_SP2_PROF
    const long long _r0 = clock64();
#endif
    for (;;) {
        const uint32_t gen = consumed / RQ_N + 1u;
        uint2 e;
        // poll with a short sleep: a tight spin competes with the entropy warps of this SM sub-partition for issue slots
        for (;;) {
            e = ld_volatile_v2(&q->e[consumed % RQ_N]);
            if (((e.x >> 27) == (gen & 31u)) && ((e.y >> 16) == (gen & 0xFFFFu))) break;
            __nanosleep(poll_ns);
        }
        P2_T(8)
        const uint32_t type = (e.x >> 24) & 7u, clr = e.x & 0xFFFFFFu;
        const int n = (int)(e.y & 0xFFFFu);
        if (type == RQ_END) break;
        for (int o = 0; o < n; o += chunk) {
            const int m = n - o < chunk ? n - o : chunk;
            const long s = di + o, idx = s + lane;
            // every predictor's inputs are loaded (5 independent shared-memory reads), the type only selects: no branches.
            //   above[i] = ring[i - X], aboveleft[i] = ring[i - X - 1];  `e` = the chunk's last pixel
            const uint32_t ab = ring.ld(idx - X), al = ring.ld(idx - X - 1);
            const uint32_t ab_e = ring.ld(s + m - 1 - X), al_e = ring.ld(s + m - 2 - X);
            const uint32_t al_s = ring.ld(s - 1 - X);
            // predictor 4: p[i] = p[i-1] + above[i] - aboveleft[i] per byte and aboveleft[i] = above[i-1], so the sum telescopes to
            // p[i] = p[s-1] + above[i] - above[s-1] -- no scan along the run
            const uint32_t g = vadd4(lastval, vsub4(ab, al_s)) & 0x00FFFFFFu, g_e = vadd4(lastval, vsub4(ab_e, al_s)) & 0x00FFFFFFu;
            uint32_t v = clr, last = clr;
            v = type == 1 ? lastval : v;  last = type == 1 ? lastval : last;
            v = type == 2 ? ab : v;       last = type == 2 ? ab_e : last;
            v = type == 5 ? al : v;       last = type == 5 ? al_e : last;
            v = type == 4 ? g : v;        last = type == 4 ? g_e : last;
            if (lane < m && idx < end) { dst[idx] = (int32_t)v; ring.st(idx, v); }
            lastval = last;
            __syncwarp();                                  // the next chunk / run may read what other lanes just wrote
        }
        di += n;
        consumed++;
        if constexpr (HBM) __threadfence_block();          // the entropy warp reads the last pixel from HBM after it sees `done`
        if (lane == 0) st_relaxed(&q->done, consumed);
        P2_T(9)
    }
#ifdef JSP_SP2_PROF
    _p2[10] = clock64() - _r0;
#endif
    P2_FLUSH
}

// ---- warp 0: entropy decode of a coded I frame ---------------------------------------------------------------------
template <class Coder, bool HBM>
__device__ __forceinline__ void sp2_entropy_iframe(Coder &ec, Producer &pq, const SpJob &J, const RingRef<HBM> ring)
{
    const long X = J.X, end = (long)J.X * J.Y;
    const int cxshift = (J.flags & SPJ_CXSHIFT0) ? 0 : 2;
    int maskcx1 = 0xFC00, shiftcx1 = 4, shiftcx = 18;
    if (J.flags & SPJ_DIFF16) { maskcx1 = 0xFF00; shiftcx1 = 2; shiftcx = 16; }
    ec.begin_iframe(J);                                    // model reset (EntroCoders.hx:81-130 / :216-227)
    ec.decodeBegin(J.src, J.len, 1);
    int cx = 0, cx1 = 0;
    long di = 0, k = 0;
    uint32_t clr = 0;
    auto decode_rgb = [&]() -> uint32_t {                 // ScreenPressor.hx:173-183
        uint32_t px = 0;
        // Coder::kUnrollChannels: three inlined copies of the colour decoder (no back edge on the symbol chain) or one -- the rANS
        // colour decoder is large, and a hot loop that does not fit the instruction cache costs more than the back edge
#pragma unroll (Coder::kUnrollChannels ? 3 : 1)
        for (int ch = 0; ch < 3; ch++) {
            const int v = ec.decodeClr(sp_ctx_index(ec, ch, cx, cx1));
            cx1 = (cx << 6) & 0xFC0; cx = v >> cxshift;
            px += (uint32_t)v << (8 * ch);
        }
        return px;
    };
    long budget = sp_run_budget(X, J.Y);
    int ptype = 0;
    bool head = true;                                      // the first X + 1 pixels: (colour, run) pairs, no predictor types (:170-197)
    bool clr_lazy = false;                                 // clr = the last pixel written so far; fetched from R when needed
    bool ctx_from_clr = false;                             // the contexts are recomputed from clr after every run of the main part (:274-275)
    P2_DECL
#ifdef JSP_SP2_PROF
    const long long _e0 = clock64();
#endif
    // ONE loop for both parts of the frame, so that the kernel holds one copy of each decoder.  Main part (:218-286): branches
    // per run are "head?", "colour follows" and the back edge -- the rest is selects.  The run budget (a frame of zero-length
    // runs must not spin for ever) is checked at the END of a main-part run instead of before the next one: a failed coder
    // decodes nothing more, so the result is the same.
    for (;;) {
        P2_T(11)
        if (head) { if (--budget < 0) ec.fail_frame(); }
        else ptype = ec.decodeP(ptype);
        P2_T(0)
        if (ptype == 0) {
            if (clr_lazy) { pq.drain(); if constexpr (HBM) clr = (uint32_t)__ldcg(J.dst + (di - 1 < end ? di - 1 : end - 1)); else clr = ring.ld(di - 1); clr_lazy = false; P2_T(4) }
            if (ctx_from_clr) { cx1 = ((int)clr & maskcx1) >> shiftcx1; cx = (int)clr >> shiftcx; }
            clr = decode_rgb();
            P2_T(1) P2_C(7)
        }
        int n = ec.decodeN(ptype);
        P2_T(2)
        if (head) {
            if (ec.failed()) return;
            k += n;
            if (n > 0) pq.push(0u, (uint32_t)n, clr);
            di += n;
            if (k >= X + 1) {
                head = false;
                if (di >= end) break;
            }
            continue;
        }
        n = (ptype == 3 || ptype > 5) ? 0 : n;             // no such predictor in an I frame: nothing is written
        // `clr = dst[lasti]` even for an empty run of predictor 1 (:252); otherwise clr = the run's last pixel
        clr_lazy = clr_lazy | (ptype == 1) | (ptype != 0 && n > 0);
        const bool put = n > 0 && !ec.failed();
        if (pq.full()) pq.wait_room();
        pq.push_if(put, (uint32_t)ptype, (uint32_t)n, clr);
        di += n;
        ctx_from_clr = true;
        --budget;
        P2_T(3) P2_C(6)
        if (di >= end || ec.failed() || budget <= 0) break;  // round 1 checked `--budget < 0` before a run: same count
    }
    if (budget <= 0 && di < end) ec.fail_frame();
#ifdef JSP_SP2_PROF
    _p2[5] = clock64() - _e0;
#endif
    P2_FLUSH
}

// ---- job pick-up: one CTA per job; in a launch that holds both coders every SM prefers ONE of them ----------------------
// The two coders' hot loops do not fit an SM's instruction cache together: with range-coder and rANS warps interleaved on
// every SM, C3's mixed launch took 147 ms where the slower coder alone takes 104 ms (round 1 saw the same: 190 vs 143 ms).
// So the jobs of a launch form two queues -- [0, n_rc) range coder, [n_rc, n_rc + n_ans) rANS -- and a CTA takes its job from
// the queue its SM prefers (SM ids below the range coder's share of the jobs prefer the range coder), falling back to the
// other queue when its own is empty.  Returns the job index, or -1 (never happens: there is one CTA per job).
__device__ __forceinline__ int sp2_take_job(uint32_t n_rc, uint32_t n_ans, uint32_t *queue)
{
    if (n_rc == 0u || n_ans == 0u || queue == nullptr) return (int)blockIdx.x;
    __shared__ int s_job;
    if (threadIdx.x == 0) {
        uint32_t smid, nsm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        const bool prefer_rc = (unsigned long long)smid * (n_rc + n_ans) < (unsigned long long)n_rc * nsm;
        int job = -1;
        for (int attempt = 0; attempt < 2 && job < 0; attempt++) {
            const bool rc = (attempt == 0) == prefer_rc;
            const uint32_t i = atomicAdd(&queue[rc ? 0 : 1], 1u);
            if (i < (rc ? n_rc : n_ans)) job = (int)(rc ? i : n_rc + i);
        }
        s_job = job;
    }
    __syncthreads();
    return s_job;
}

// ---- coded I frames: entropy warp + reconstruction warp (see the header of this file) ----------------------------------
template <class Coder, bool HBM>
__device__ __forceinline__ void sp2_iframe_body(const SpJob &J, Coder &ec, RunQueue &rq, uint32_t *ring)
{
    const RingRef<HBM> rr = HBM ? RingRef<HBM>{reinterpret_cast<uint32_t *>(J.dst), 0xFFFFFFFFu} : RingRef<HBM>{ring, sp_ring_size(J.X) - 1u};
    const int lane = threadIdx.x & 31;
    const int warp = sp2_pick_roles(&rq);                  // 0 = entropy, 1 = reconstruction
    bool failed = false;
    if (warp == 0) {
        ec.open_iframe(J);
        Producer pq{&rq, 0u, 0u};
        sp2_entropy_iframe(ec, pq, J, rr);
        pq.push(RQ_END, 0u, 0u);
        failed = ec.failed();
    } else {
        sp2_recon_iframe(&rq, J, rr);
    }
    __syncthreads();                                       // R has written every pixel E queued
    if (warp == 0) {
        ec.close_frame(J);
        uint32_t bits = ST_CHANGED;
        if (failed) { bits = ST_ERROR; sp_undo_frame(J, true); }
        if (lane == 0) { atomicOr(J.status, bits); if (J.symbols) *J.symbols = ec.nsym; }
    }
    if (J.done) {                                          // the host may copy the picture out while the launch still runs
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t *>(J.done) = 1u;
    }
}

constexpr uint32_t SP2_ANS_I_BYTES = ANS_SMALL_I_BYTES + (uint32_t)sizeof(AnsWork);
constexpr uint32_t SP2_I_BYTES = RC_SHARED_I_BYTES > SP2_ANS_I_BYTES ? RC_SHARED_I_BYTES : SP2_ANS_I_BYTES;

// CODERS: 1 = the launch holds range-coder frames only, 2 = rANS only, 3 = both (a kernel that holds one coder keeps that
// coder's registers and code layout: the range coder is 7 % slower per symbol inside the two-coder kernel)
template <bool HBM, int CODERS>
__global__ void __launch_bounds__(64)
sp2_i_kernel(const SpJob *__restrict__ jobs, uint32_t n_rc, uint32_t n_ans, uint32_t *queue, uint32_t ring_words)
{
    constexpr uint32_t BYTES = CODERS == 1 ? RC_SHARED_I_BYTES : (CODERS == 2 ? SP2_ANS_I_BYTES : SP2_I_BYTES);
    __shared__ alignas(16) uint8_t shm[BYTES];             // one coder's tables: without those only P frames use
    __shared__ RunQueue rq;
    extern __shared__ uint32_t ring[];                     // the last X + 1 pixels (power of two > X + 65 words)
    const int job = sp2_take_job(n_rc, n_ans, queue);
    if (job < 0) return;
    const SpJob J = jobs[job];
    if (CODERS == 2 || (CODERS == 3 && (J.flags & SPJ_ANS))) {
        if constexpr (CODERS != 1) {
            AnsCoder ec;
            ec.small = reinterpret_cast<AnsSmall *>(shm); ec.wk = reinterpret_cast<AnsWork *>(shm + ANS_SMALL_I_BYTES); ec.small_bytes = ANS_SMALL_I_BYTES;
            sp2_iframe_body<AnsCoder, HBM>(J, ec, rq, ring);
        }
    } else {
        if constexpr (CODERS != 2) {
            RcCoder ec;
            ec.bind(reinterpret_cast<RcShared *>(shm), RC_SMALL_I_BYTES);
            sp2_iframe_body<RcCoder, HBM>(J, ec, rq, ring);
        }
    }
}

// ---- P frames and model resets of flat frames: one warp, round 1's frame loop (sp_common.cuh) on the new symbol decoders ----
template <class Coder>
__device__ __forceinline__ void sp2_pframe_body(const SpJob &J, Coder &ec, uint32_t *ptile)
{
    ec.open_pframe(J);
    uint32_t bits = 0;
    if (J.flags & SPJ_RENEW) ec.renewI();
    else sp_decode_pframe(ec, J, bits, ptile);
    ec.close_frame(J);
    if (ec.failed()) {
        bits = ST_ERROR;
        if (!(J.flags & SPJ_RENEW)) sp_undo_frame(J, false);
    }
    __syncwarp();
    if (lane_id() == 0) { if (bits) atomicOr(J.status, bits); if (J.symbols) *J.symbols = ec.nsym; }
    sp_signal_done(J);
}

constexpr uint32_t SP2_P_BYTES = sizeof(RcShared) > sizeof(AnsShared) ? (uint32_t)sizeof(RcShared) : (uint32_t)sizeof(AnsShared);

template <int CODERS>
__global__ void __launch_bounds__(32)
sp2_p_kernel(const SpJob *__restrict__ jobs, uint32_t n_rc, uint32_t n_ans, uint32_t *queue, uint32_t tile_words)
{
    constexpr uint32_t BYTES = CODERS == 1 ? (uint32_t)sizeof(RcShared) : (CODERS == 2 ? (uint32_t)sizeof(AnsShared) : SP2_P_BYTES);
    __shared__ alignas(16) uint8_t shm[BYTES];
    extern __shared__ uint32_t ptile_mem[];
    const int job = sp2_take_job(n_rc, n_ans, queue);
    if (job < 0) return;
    const SpJob J = jobs[job];
    uint32_t *ptile = tile_words >= SP_PTILE_WORDS ? ptile_mem : nullptr;
    if (CODERS == 2 || (CODERS == 3 && (J.flags & SPJ_ANS))) {
        if constexpr (CODERS != 1) {
            AnsCoder ec;
            AnsShared *sh = reinterpret_cast<AnsShared *>(shm);
            ec.small = &sh->small; ec.wk = &sh->work; ec.small_bytes = (uint32_t)sizeof(AnsSmall);
            sp2_pframe_body(J, ec, ptile);
        }
    } else {
        if constexpr (CODERS != 2) {
            RcCoder ec;
            ec.bind(reinterpret_cast<RcShared *>(shm), (uint32_t)sizeof(RcSmall));
            sp2_pframe_body(J, ec, ptile);
        }
    }
}

}  // namespace g2

#ifdef JSP_SP2_PROF
extern "C" __attribute__((visibility("default"))) int jsp_debug_sp2_profile(unsigned long long *out, int reset)
{
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out, g2::g_sp2_prof, sizeof z) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(g2::g_sp2_prof, z, sizeof z);
    return 0;
}
#endif

#ifdef JSP_PROFILE_SECTIONS
// the colour-decoder section timers of sp_ans.cuh as THIS translation unit's kernels accumulated them
extern "C" __attribute__((visibility("default"))) int jsp_debug_ans2_profile(unsigned long long *out, int reset)
{
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out, g_ans_prof, sizeof z) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(g_ans_prof, z, sizeof z);
    return 0;
}
#endif

// ---- host side -----------------------------------------------------------------------------------------------------
int sp_generation()
{
    static const int gen = [] { const char *e = getenv("JSP_SP_GEN"); return e && e[0] == '1' ? 1 : 2; }();
    return gen;
}

size_t sp2_rc_state_bytes() { return (sizeof(g2::RcState) + 255) & ~(size_t)255; }
void sp2_rc_state_init(void *d_state, void *d_rows, uint32_t gen0, cudaStream_t st)
{
    g2::RcState h;
    memset(&h, 0, sizeof h);
    h.gen = gen0; h.rows = reinterpret_cast<uint32_t *>(d_rows);
    cudaStreamSynchronize(st);      // ordered after the memsets queued on st
    cudaMemcpy(reinterpret_cast<char *>(d_state) + offsetof(g2::RcState, gen), &h.gen, sizeof(g2::RcState) - offsetof(g2::RcState, gen),
               cudaMemcpyHostToDevice);
}

namespace {
struct DevAux {                                            // per device: side streams so that the kernels of one level overlap
    cudaStream_t s[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
    bool attr = false, ok = false;
};
DevAux g_aux[64];
std::atomic<unsigned long long> g_aux_ready{0};

DevAux *aux_for_current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    DevAux &A = g_aux[dev];
    const unsigned long long bit = 1ull << dev;
    if (!(g_aux_ready.load(std::memory_order_acquire) & bit)) {
        // one host thread drives a device (jsp_batch_decode: one thread per GPU), so no lock is needed per entry
        bool ok = true;
        for (int i = 0; i < 3; i++) {
            ok = ok && cudaStreamCreateWithFlags(&A.s[i], cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&A.join[i], cudaEventDisableTiming) == cudaSuccess;
        }
        ok = ok && cudaEventCreateWithFlags(&A.fork, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(g2::sp2_i_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(g2::sp2_i_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(g2::sp2_i_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess;
        if (const char *e = getenv("JSP_SP2_POLL_NS")) {
            const unsigned int ns = (unsigned int)atoi(e);
            ok = ok && cudaMemcpyToSymbol(g2::g_sp2_poll_ns, &ns, sizeof ns) == cudaSuccess;
        }
        A.ok = ok;
        g_aux_ready.fetch_or(bit, std::memory_order_release);
    }
    return A.ok ? &A : nullptr;
}
}  // namespace

// Jobs of one dependency level, ordered by the planner: [range-coder I frames | rANS I frames | range-coder P frames and
// model resets | rANS P frames and model resets].  I frames and P frames are two concurrent launches (the level lasts as long
// as its slowest frame); d_queue = four zeroed counters (two per launch) for the per-SM coder preference.
bool launch_sp2_level(const SpJob *d_jobs, uint32_t n_rc_i, uint32_t n_ans_i, uint32_t n_rc_p, uint32_t n_ans_p, uint32_t max_width,
                      uint32_t *d_queue, cudaStream_t st)
{
    DevAux *A = aux_for_current_device();
    uint32_t words = 1024;                                             // at least the P-frame block tile (SP_PTILE_WORDS)
    while (words <= max_width + 65u) words <<= 1;                      // >= sp_ring_size(max_width)
    if (!A) return false;                                              // caller falls back to the first-generation kernel
    if (words > 16384u) words = 0;                                     // > 64 KB (pictures wider than 16 318): the I-frame kernel
                                                                       // reads the row above from the picture in HBM instead
    const uint32_t n_i = n_rc_i + n_ans_i, n_p = n_rc_p + n_ans_p;
    cudaStream_t sp = st;
    const bool fork = n_i && n_p;
    if (fork) { cudaEventRecord(A->fork, st); sp = A->s[0]; cudaStreamWaitEvent(sp, A->fork, 0); }
    if (n_i) {
        const int coders = (n_rc_i ? 1 : 0) | (n_ans_i ? 2 : 0);
        const size_t dyn = (size_t)words * 4;
#define JSP_LAUNCH_I(HBM, C) g2::sp2_i_kernel<HBM, C><<<n_i, 64, dyn, st>>>(d_jobs, n_rc_i, n_ans_i, d_queue, words)
        if (words) { if (coders == 1) JSP_LAUNCH_I(false, 1); else if (coders == 2) JSP_LAUNCH_I(false, 2); else JSP_LAUNCH_I(false, 3); }
        else       { if (coders == 1) JSP_LAUNCH_I(true, 1);  else if (coders == 2) JSP_LAUNCH_I(true, 2);  else JSP_LAUNCH_I(true, 3); }
#undef JSP_LAUNCH_I
    }
    if (n_p) {
        const int coders = (n_rc_p ? 1 : 0) | (n_ans_p ? 2 : 0);
        uint32_t *q = d_queue ? d_queue + 2 : nullptr;
        if (coders == 1) g2::sp2_p_kernel<1><<<n_p, 32, (size_t)1024 * 4, sp>>>(d_jobs + n_i, n_rc_p, n_ans_p, q, 1024);
        else if (coders == 2) g2::sp2_p_kernel<2><<<n_p, 32, (size_t)1024 * 4, sp>>>(d_jobs + n_i, n_rc_p, n_ans_p, q, 1024);
        else g2::sp2_p_kernel<3><<<n_p, 32, (size_t)1024 * 4, sp>>>(d_jobs + n_i, n_rc_p, n_ans_p, q, 1024);
    }
    if (fork) { cudaEventRecord(A->join[0], sp); cudaStreamWaitEvent(st, A->join[0], 0); }
    return true;
}

}  // namespace jsp
