// sp_decode.cu -- the ScreenPressor entropy-decode kernel: one warp (one 32-thread CTA) per frame of one stream.
// Range-coder (v2, sp_rc.cuh) and rANS (v3 / v4, sp_ans.cuh) jobs share ONE launch, so that the streams of a level run
// concurrently whatever their coder is (two launches would each wait for their own slowest warp).
// Replaces the per-frame entry of reference src/ScreenPressor.hx (DecompressI :117-295, DecompressP :302-484).
#include "sp_rc.cuh"
#include <atomic>
#include "sp_ans.cuh"

namespace jsp {
namespace {

constexpr size_t SP_SMEM = sizeof(AnsShared) > sizeof(RcShared) ? sizeof(AnsShared) : sizeof(RcShared);

__global__ void __launch_bounds__(32)
sp_decode_kernel(const SpJob *__restrict__ jobs, uint32_t ring_words, uint32_t n_rc, uint32_t *__restrict__ queue)
{
    __shared__ alignas(16) uint8_t smem[SP_SMEM];
    extern __shared__ uint32_t ring_mem[];          // the last X + 1 pixels of an I frame (sp_segment_ring)
    uint32_t idx = blockIdx.x;
    if (queue) {
        // Both coders in one launch: jobs [0, n_rc) are range-coder frames, the rest rANS frames.  The warps that share
        // an SM share its instruction cache, and the two decoders together do not fit it -- so an SM prefers one
        // coder: the low SM ids take from the range-coder queue, the others from the rANS queue, each falling back to
        // the other queue once its own is empty (gridDim.x == number of jobs, so every warp gets exactly one).
        if ((threadIdx.x & 31) == 0) {
            uint32_t smid, nsm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
            const uint32_t n = gridDim.x, n_ans = n - n_rc;
            const bool want_rc = (uint64_t)smid * n < (uint64_t)n_rc * nsm;
            uint32_t t = atomicAdd(&queue[want_rc ? 0 : 1], 1u);
            if (want_rc) idx = t < n_rc ? t : n_rc + atomicAdd(&queue[1], 1u);
            else idx = t < n_ans ? n_rc + t : atomicAdd(&queue[0], 1u);
        }
        idx = __shfl_sync(0xffffffffu, idx, 0);
    }
    const SpJob J = jobs[idx];
    uint32_t *ring = sp_ring_size(J.X) <= ring_words ? ring_mem : nullptr;     // pictures too wide for it read global memory
    uint32_t *ptile = ring_words >= SP_PTILE_WORDS ? ring_mem : nullptr;       // P frames use the same memory for their block tile
#ifdef JSP_NO_PTILE
    ptile = nullptr;
#endif
    if (J.flags & SPJ_ANS) sp_ans_run(J, *reinterpret_cast<AnsShared *>(smem), ring, ptile);
    else sp_rc_run(J, *reinterpret_cast<RcShared *>(smem), ring, ptile);
}

}  // namespace

size_t sp_rc_state_bytes() { return (sizeof(RcState) + 255) & ~(size_t)255; }
size_t sp_rc_rows_bytes() { return (size_t)RC_ROWS * RC_ROW_STRIDE * 4; }

// host-side init of one stream's state: generation gen0, tables irrelevant until the first I frame
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

size_t sp_ans_state_bytes() { return (sizeof(AnsState) + 255) & ~(size_t)255; }
size_t sp_ans_ctx_bytes() { return (size_t)ANS_NCTX * (ANS_HDR_BYTES + ANS_BODY_BYTES); }

// host-side init of one stream's state: generation gen0 (headers are zeroed: generation 0 = "no context")
void sp_ans_state_init(void *d_state, void *d_ctx, uint32_t gen0, cudaStream_t st)
{
    struct { uint32_t gen; uint32_t pad[3]; uint4 *hdrs; uint4 *bodies; } h;
    memset(&h, 0, sizeof h);
    h.gen = gen0;
    h.hdrs = reinterpret_cast<uint4 *>(d_ctx);
    h.bodies = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(d_ctx) + (size_t)ANS_NCTX * ANS_HDR_BYTES);
    static_assert(sizeof(h) == sizeof(AnsState) - offsetof(AnsState, gen), "AnsState tail layout");
    cudaStreamSynchronize(st);
    cudaMemcpy(reinterpret_cast<char *>(d_state) + offsetof(AnsState, gen), &h, sizeof h, cudaMemcpyHostToDevice);
}

#ifdef JSP_PROFILE_SECTIONS
extern "C" __attribute__((visibility("default"))) int jsp_debug_sp_profile(unsigned long long *out, int reset)
{
    unsigned long long z[8] = {0};
    if (cudaMemcpyFromSymbol(out, g_sp_prof, sizeof z) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(g_sp_prof, z, sizeof z);
    return 0;
}
extern "C" __attribute__((visibility("default"))) int jsp_debug_spp_profile(unsigned long long *out, int reset)
{
    unsigned long long z[10] = {0};
    if (cudaMemcpyFromSymbol(out, g_spp_prof, sizeof z) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(g_spp_prof, z, sizeof z);
    return 0;
}
extern "C" __attribute__((visibility("default"))) int jsp_debug_ans_profile(unsigned long long *out, int reset)
{
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out, g_ans_prof, sizeof z) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(g_ans_prof, z, sizeof z);
    return 0;
}
extern "C" __attribute__((visibility("default"))) int jsp_debug_rc_profile(unsigned long long *out, int reset)
{
    unsigned long long z[8] = {0};
    if (cudaMemcpyFromSymbol(out, g_rc_prof, sizeof z) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(g_rc_prof, z, sizeof z);
    return 0;
}
#endif

void launch_sp_decode(const SpJob *d_jobs, uint32_t n_jobs, uint32_t max_width, uint32_t n_rc, uint32_t *d_queue, cudaStream_t st)
{
    if (!n_jobs) return;
    uint32_t words = 1024;                                             // at least the P-frame block tile (SP_PTILE_WORDS)
    while (words <= max_width + 65u) words <<= 1;                    // >= sp_ring_size(max_width)
    if (words > 16384u) words = 1024;                                  // > 64 KB: such I frames fall back to global reads
    static_assert(SP_PTILE_WORDS <= 1024, "P-frame tile fits the minimum dynamic shared memory");
    // the opt-in is per device (context), not per process: jsp_batch_decode(n_gpus > 1) launches from one thread per device
    static std::atomic<unsigned long long> attr_devices{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_devices.load(std::memory_order_acquire) & bit)) {
        if (cudaFuncSetAttribute(sp_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) == cudaSuccess)
            attr_devices.fetch_or(bit, std::memory_order_release);
    }
    const bool mixed = n_rc != 0 && n_rc != n_jobs && d_queue;
    sp_decode_kernel<<<n_jobs, 32, (size_t)words * 4, st>>>(d_jobs, words, n_rc, mixed ? d_queue : nullptr);
}

}  // namespace jsp
