// frame_ops.cu -- wide whole-picture kernels: copy / fill (unchanged frames, ScreenPressor flat frames and
// P-frame pre-copies) and the significance compare of MSVideo1.hx:195-204.  Pure HBM streaming.
#include "common.cuh"

namespace jsp {
namespace {

constexpr int COPY_THREADS = 256;
constexpr int COPY_UNROLL = 4;

// grid = (chunks, jobs): every CTA streams COPY_THREADS*COPY_UNROLL 16-byte units per iteration
__global__ void __launch_bounds__(COPY_THREADS)
frame_copy_kernel(const CopyJob *__restrict__ jobs)
{
    const CopyJob J = jobs[blockIdx.y];
    uint4 *__restrict__ d = reinterpret_cast<uint4 *>(J.dst);
    const uint32_t n = J.n_vec4;
    const uint32_t stride = gridDim.x * COPY_THREADS * COPY_UNROLL;
    if (J.src) {
        const uint4 *__restrict__ s = reinterpret_cast<const uint4 *>(J.src);
        for (uint32_t base = blockIdx.x * COPY_THREADS * COPY_UNROLL + threadIdx.x; base < n; base += stride) {
            uint4 v[COPY_UNROLL];
#pragma unroll
            for (int k = 0; k < COPY_UNROLL; k++) {
                const uint32_t i = base + k * COPY_THREADS;
                if (i < n) v[k] = __ldcs(s + i);
            }
#pragma unroll
            for (int k = 0; k < COPY_UNROLL; k++) {
                const uint32_t i = base + k * COPY_THREADS;
                if (i < n) __stcs(d + i, v[k]);
            }
        }
    } else {
        const uint4 v = make_uint4(J.value, J.value, J.value, J.value);
        for (uint32_t base = blockIdx.x * COPY_THREADS * COPY_UNROLL + threadIdx.x; base < n; base += stride) {
#pragma unroll
            for (int k = 0; k < COPY_UNROLL; k++) {
                const uint32_t i = base + k * COPY_THREADS;
                if (i < n) __stcs(d + i, v);
            }
        }
    }
}

// MSVideo1.hx:195-204: does any pixel from `first_px` on differ from the previous picture?
// KEY: the same compare for key frames as Manager.frames_differ_significantly does it (Manager.hx:413-419)
template <bool KEY>
__global__ void __launch_bounds__(256)
signif_kernel(const int32_t *const *__restrict__ cur, const int32_t *const *__restrict__ prev,
              uint32_t *const *__restrict__ status, const uint32_t *__restrict__ first_px,
              const uint32_t *__restrict__ npx)
{
    const uint32_t j = blockIdx.y;
    // only frames that pass the block-row test and had a previous picture reach the pixel compare
    const uint32_t stv = *status[j];
    if (!KEY && (stv & (ST_SIGNIF_ROWS | ST_HAS_PREV)) != (ST_SIGNIF_ROWS | ST_HAS_PREV)) return;
    const int32_t *c = cur[j], *p = prev[j];
    const uint32_t n = npx[j];
    bool diff = false;
    for (uint32_t i = first_px[j] + blockIdx.x * 256 + threadIdx.x; i < n && !diff; i += gridDim.x * 256)
        diff = c[i] != p[i];
    if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(status[j], KEY ? ST_KEYDIFF : ST_PIXDIFF);
}

// Manager.fill_bitmap_data (canvas branch, Manager.hx:363-381): 0x00RRGGBB -> Int32 view of canvas bytes R,G,B,A.
// grid = (chunks, jobs); one thread converts 4 horizontally adjacent pixels (16-byte load / store when X % 4 == 0).
__global__ void __launch_bounds__(256)
display_kernel(const DisplayJob *__restrict__ jobs)
{
    const DisplayJob J = jobs[blockIdx.y];
    const uint32_t X = J.X, Y = J.Y;
    auto conv = [&](uint32_t c) -> uint32_t {
        return J.from_rgb15 ? (0xFF000000u | (c << 3)) : (0xFF000000u | __byte_perm(c, 0, 0x4012));   // bytes (B,G,R,x) -> (R,G,B,x)
    };
    if ((X & 3u) == 0) {
        const uint32_t xv = X >> 2, n = xv * Y;
        for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
            const uint32_t y = i / xv, x4 = i - y * xv;
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(J.src + (size_t)y * X) + x4);
            const uint32_t yo = J.flip ? Y - 1 - y : y;
            __stcs(reinterpret_cast<uint4 *>(J.dst + (size_t)yo * X) + x4, make_uint4(conv(v.x), conv(v.y), conv(v.z), conv(v.w)));
        }
    } else {
        const uint32_t n = X * Y;
        for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
            const uint32_t y = i / X, x = i - y * X;
            const uint32_t yo = J.flip ? Y - 1 - y : y;
            J.dst[(size_t)yo * X + x] = (int32_t)conv((uint32_t)J.src[i]);
        }
    }
}

// prevFrame is non-null for frame f <=> some earlier frame of the stream altered pixels
// (MSVideo1.hx:206-208: `if (changes) prevFrame = dst`).  One thread per stream, frames in order.
__global__ void status_scan_kernel(uint32_t *__restrict__ status, const uint32_t *__restrict__ stream_first,
                                   const uint32_t *__restrict__ stream_count, uint32_t n_streams, int init_has_prev)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    uint32_t *st = status + stream_first[s];
    bool has_prev = init_has_prev != 0;
    for (uint32_t f = 0; f < stream_count[s]; f++) {
        uint32_t v = st[f];
        if (has_prev) { v |= ST_HAS_PREV; st[f] = v; }
        if (v & ST_CHANGED) has_prev = true;
    }
}

// significant_changes per codec: RGB555 MSVideo1.hx:187-204; 8-bit :372-388 (the pixel loop never runs in
// JavaScript because insign_lines is undefined there, so a previous picture forces `false`);
// ScreenPressor sets ST_SIGNIFICANT itself (ScreenPressor.hx:347-352).
__global__ void status_final_kernel(uint32_t *__restrict__ status, const uint8_t *__restrict__ frame_codec,
                                    uint32_t n, int exact)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v = status[i];
    const uint32_t codec = frame_codec[i];
    bool sig = (v & ST_SIGNIFICANT) != 0;
    if (codec == 1)      sig = (v & ST_SIGNIF_ROWS) && (!(v & ST_HAS_PREV) || !exact || (v & ST_PIXDIFF));
    else if (codec == 2) sig = (v & ST_SIGNIF_ROWS) && !(v & ST_HAS_PREV);
    status[i] = sig ? (v | ST_SIGNIFICANT) : (v & ~ST_SIGNIFICANT);
}

}  // namespace

void launch_status_scan(uint32_t *d_status, const uint32_t *d_stream_first, const uint32_t *d_stream_count,
                        uint32_t n_streams, int init_has_prev, cudaStream_t st)
{
    if (n_streams) status_scan_kernel<<<(n_streams + 127) / 128, 128, 0, st>>>(d_status, d_stream_first, d_stream_count, n_streams, init_has_prev);
}

void launch_status_final(uint32_t *d_status, const uint8_t *d_frame_codec, uint32_t n, int exact, cudaStream_t st)
{
    if (n) status_final_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_status, d_frame_codec, n, exact);
}

void launch_frame_copy(const CopyJob *d_jobs, uint32_t n_jobs, uint32_t max_vec4, int sm_count, cudaStream_t st)
{
    if (n_jobs == 0 || max_vec4 == 0) return;
    const uint32_t per_cta = COPY_THREADS * COPY_UNROLL;
    uint32_t chunks = (max_vec4 + per_cta - 1) / per_cta;
    // enough CTAs to fill the machine a few times over, no more (grid-stride covers the rest)
    const uint32_t want = (uint32_t)sm_count * 8u;
    const uint32_t cap = (want + n_jobs - 1) / n_jobs;
    if (chunks > cap) chunks = cap;
    if (chunks == 0) chunks = 1;
    for (uint32_t j0 = 0; j0 < n_jobs; j0 += 65535u) {
        const uint32_t nj = n_jobs - j0 < 65535u ? n_jobs - j0 : 65535u;
        frame_copy_kernel<<<dim3(chunks, nj), COPY_THREADS, 0, st>>>(d_jobs + j0);
    }
}

void launch_display(const DisplayJob *d_jobs, uint32_t n_jobs, uint32_t max_pixels, int sm_count, cudaStream_t st)
{
    if (n_jobs == 0 || max_pixels == 0) return;
    uint32_t chunks = (max_pixels / 4 + 255) / 256;
    const uint32_t cap = ((uint32_t)sm_count * 16u + n_jobs - 1) / n_jobs;
    if (chunks > cap) chunks = cap;
    if (chunks == 0) chunks = 1;
    for (uint32_t j0 = 0; j0 < n_jobs; j0 += 65535u) {
        const uint32_t nj = n_jobs - j0 < 65535u ? n_jobs - j0 : 65535u;
        display_kernel<<<dim3(chunks, nj), 256, 0, st>>>(d_jobs + j0);
    }
}

void launch_signif(const int32_t *const *d_cur, const int32_t *const *d_prev, uint32_t *const *d_status,
                   const uint32_t *d_first_px, const uint32_t *d_npx, uint32_t n_jobs, int sm_count, cudaStream_t st,
                   bool key_frames)
{
    if (n_jobs == 0) return;
    uint32_t chunks = ((uint32_t)sm_count * 8u + n_jobs - 1) / n_jobs;
    if (chunks == 0) chunks = 1;
    for (uint32_t j0 = 0; j0 < n_jobs; j0 += 65535u) {
        const uint32_t nj = n_jobs - j0 < 65535u ? n_jobs - j0 : 65535u;
        if (key_frames) signif_kernel<true><<<dim3(chunks, nj), 256, 0, st>>>(d_cur + j0, d_prev + j0, d_status + j0, d_first_px + j0, d_npx + j0);
        else signif_kernel<false><<<dim3(chunks, nj), 256, 0, st>>>(d_cur + j0, d_prev + j0, d_status + j0, d_first_px + j0, d_npx + j0);
    }
}

}  // namespace jsp
