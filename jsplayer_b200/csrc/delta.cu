// delta.cu -- opt-in end-to-end path for hosts that keep ONE picture per stream and update it in place.
//
// The IVideoCodec contract returns a whole Int32Array picture per frame (IVideoCodec.hx:11-29), 4 bytes per pixel over PCIe,
// which bounds every end-to-end number of this library.  But the consumer the reference drives -- a
// player -- shows frames in order and keeps the previous picture (Manager.hx:470-477, PreviousFrame), and inter-frame screen
// content changes a few percent of its blocks per frame (ScreenPressor.hx:336-352: block type 0 = unchanged; MSVideo1 skip
// runs, MSVideo1.hx:131-133).  So after the batch is decoded on the device, a wide kernel compares every picture with its
// stream's previous one in 16x16 blocks and packs the blocks that differ (coordinates + 256 pixels); only those cross PCIe,
// and host threads patch them into the stream's picture, frame by frame in order.  Exact by construction: a block is sent iff
// some pixel of it differs.  A stream's first frame sends every block.  HBM-bound: 8 bytes per pixel read, once.
#include "batch.cuh"
#include <algorithm>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <vector>
#include <cstring>

namespace jsp {

struct DeltaJob {
    const int32_t *cur;
    const int32_t *prev;      // nullptr: every block is sent
    uint32_t X, Y, nbx, nb;   // picture size; 16x16 blocks per row / in total
    uint32_t index;           // job index within the unit (the host maps it to stream / frame)
    uint32_t pad;
};
struct DeltaEntry { uint32_t job, block; };
struct DeltaStage {
    DeltaJob *d_jobs = nullptr, *h_jobs = nullptr; DeltaEntry *d_ent = nullptr, *h_ent = nullptr; int32_t *d_pix = nullptr, *h_pix = nullptr;
    uint32_t *d_cnt = nullptr, *h_cnt = nullptr; cudaEvent_t packed = nullptr, copied = nullptr;
};
struct DeltaStages { DeltaStage st[2]; size_t blocks = 0, jobs = 0; };
static void delta_free(DeltaStages &DS)
{
    for (DeltaStage &S : DS.st) {
        if (S.d_jobs) cudaFree(S.d_jobs); if (S.h_jobs) cudaFreeHost(S.h_jobs);
        if (S.d_ent) cudaFree(S.d_ent);   if (S.h_ent) cudaFreeHost(S.h_ent);
        if (S.d_pix) cudaFree(S.d_pix);   if (S.h_pix) cudaFreeHost(S.h_pix);
        if (S.d_cnt) cudaFree(S.d_cnt);   if (S.h_cnt) cudaFreeHost(S.h_cnt);
        if (S.packed) cudaEventDestroy(S.packed); if (S.copied) cudaEventDestroy(S.copied);
        S = DeltaStage{};
    }
    DS.blocks = DS.jobs = 0;
}
void delta_release(jsp_batch *b) { if (b->delta) { delta_free(*b->delta); delete b->delta; b->delta = nullptr; } }

namespace {
constexpr int DELTA_WARPS = 8;

// one warp per 16x16 block: lane = (row pair, half row): 8 pixels of two rows each
__global__ void __launch_bounds__(DELTA_WARPS * 32)
delta_pack_kernel(const DeltaJob *__restrict__ jobs, DeltaEntry *__restrict__ entries, int32_t *__restrict__ pixels, uint32_t *__restrict__ counter, uint32_t cap)
{
    const DeltaJob J = jobs[blockIdx.y];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t blk = blockIdx.x * DELTA_WARPS + warp;
    if (blk >= J.nb) return;
    const uint32_t by = blk / J.nbx, bx = blk - by * J.nbx;
    const uint32_t x0 = bx * 16 + (lane & 1) * 8;
    const bool vec = (J.X & 3u) == 0 && ((reinterpret_cast<uintptr_t>(J.cur) | reinterpret_cast<uintptr_t>(J.prev)) & 15u) == 0;
    uint32_t c[16];
    bool diff = J.prev == nullptr;
#pragma unroll
    for (int h = 0; h < 2; h++) {                       // rows lane/2 and lane/2 + ... : 16 rows = 16 lane pairs x 1 row; two passes of 8 px
        const uint32_t y = by * 16 + (lane >> 1);
        const uint32_t x = x0 + h * 4;
        const size_t off = (size_t)y * J.X + x;
        uint32_t p[4] = {0, 0, 0, 0};
        if (y < J.Y && vec && x + 4 <= J.X) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(J.cur + off));
            c[4 * h] = v.x; c[4 * h + 1] = v.y; c[4 * h + 2] = v.z; c[4 * h + 3] = v.w;
            if (J.prev) { const uint4 w = __ldcs(reinterpret_cast<const uint4 *>(J.prev + off)); p[0] = w.x; p[1] = w.y; p[2] = w.z; p[3] = w.w; }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool in = y < J.Y && x + k < J.X;
                c[4 * h + k] = in ? (uint32_t)J.cur[off + k] : 0u;
                p[k] = (in && J.prev) ? (uint32_t)J.prev[off + k] : c[4 * h + k];
            }
        }
        if (J.prev) diff = diff || c[4 * h] != p[0] || c[4 * h + 1] != p[1] || c[4 * h + 2] != p[2] || c[4 * h + 3] != p[3];
    }
    if (!__any_sync(0xffffffffu, diff)) return;
    uint32_t slot = 0;
    if (lane == 0) slot = atomicAdd(counter, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= cap) return;                            // cannot happen: cap = all blocks of the unit
    if (lane == 0) entries[slot] = DeltaEntry{J.index, blk};
    // packed block: 16 rows x 16 pixels, row-major; this lane holds row lane/2, pixels (lane&1)*8 .. +8
    uint4 *dst = reinterpret_cast<uint4 *>(pixels + (size_t)slot * 256 + (lane >> 1) * 16 + (lane & 1) * 8);
    dst[0] = make_uint4(c[0], c[1], c[2], c[3]);
    dst[1] = make_uint4(c[4], c[5], c[6], c[7]);
}
}  // namespace

}  // namespace jsp

using namespace jsp;

extern "C" __attribute__((visibility("default")))
uint64_t jsp_batch_delta_bytes(jsp_batch *b) { return b ? b->delta_d2h_bytes : 0; }

extern "C" __attribute__((visibility("default")))
int jsp_batch_decode_host_delta(jsp_batch *b, int32_t *const *stream_pictures, uint8_t *flags, jsp_frame_fn on_frame, void *user)
{
    if (!b || !stream_pictures) { set_error("jsp_batch_decode_host_delta: bad arguments"); return -1; }
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    if (jsp_batch_upload(b) || jsp_batch_run(b)) return -1;
    std::vector<uint8_t> fl(b->frames.size());
    if (jsp_batch_results(b, fl.data())) return -1;           // (may re-plan and decode again: see jsp_batch_results)
    if (flags) memcpy(flags, fl.data(), fl.size());

    const int ns = (int)b->streams.size();
    int max_frames = 0;
    for (const StreamRec &S : b->streams) max_frames = std::max(max_frames, S.n_frames);
    // units: (frame index k, a run of streams) whose worst case -- every block differs -- fits the staging buffers
    const size_t STAGE_BLOCKS = (size_t)1 << 19;               // 512 Ki blocks = 512 MiB of pixels per staging buffer
    struct Unit { int k, s_lo, s_hi; };
    std::vector<Unit> units;
    size_t worst_unit = 0, max_jobs = 0;
    for (int k = 0; k < max_frames; k++) {
        int s = 0;
        while (s < ns) {
            size_t blocks = 0; int s_lo = s;
            while (s < ns) {
                const StreamRec &S = b->streams[s];
                const size_t nb = S.n_frames > k ? (size_t)((S.w + 15) / 16) * ((S.h + 15) / 16) : 0;
                if (blocks && blocks + nb > STAGE_BLOCKS) break;
                blocks += nb; s++;
            }
            units.push_back({k, s_lo, s});
            worst_unit = std::max(worst_unit, blocks);
            max_jobs = std::max(max_jobs, (size_t)(s - s_lo));
        }
    }
    b->delta_d2h_bytes = 0;
    if (units.empty()) return 0;
    // two staging sets (device + pinned host), kept with the batch: unit u packs while unit u-1 crosses PCIe and is patched in
    if (!b->delta) b->delta = new DeltaStages();
    DeltaStages &DS = *b->delta;
    bool ok = true;
    if (DS.blocks < worst_unit || DS.jobs < max_jobs) {
        delta_free(DS);
        for (DeltaStage &S : DS.st) {
            ok = ok && JSP_CUDA(cudaMalloc((void **)&S.d_jobs, max_jobs * sizeof(DeltaJob))) && JSP_CUDA(cudaHostAlloc((void **)&S.h_jobs, max_jobs * sizeof(DeltaJob), cudaHostAllocDefault));
            ok = ok && JSP_CUDA(cudaMalloc((void **)&S.d_ent, worst_unit * sizeof(DeltaEntry))) && JSP_CUDA(cudaHostAlloc((void **)&S.h_ent, worst_unit * sizeof(DeltaEntry), cudaHostAllocDefault));
            ok = ok && JSP_CUDA(cudaMalloc((void **)&S.d_pix, worst_unit * 1024)) && JSP_CUDA(cudaHostAlloc((void **)&S.h_pix, worst_unit * 1024, cudaHostAllocDefault));
            ok = ok && JSP_CUDA(cudaMalloc((void **)&S.d_cnt, 4)) && JSP_CUDA(cudaHostAlloc((void **)&S.h_cnt, 4, cudaHostAllocDefault));
            ok = ok && JSP_CUDA(cudaEventCreateWithFlags(&S.packed, cudaEventDisableTiming)) && JSP_CUDA(cudaEventCreateWithFlags(&S.copied, cudaEventDisableTiming));
        }
        if (!ok) { delta_free(DS); return -1; }
        DS.blocks = worst_unit; DS.jobs = max_jobs;
    }
    DeltaStage *st = DS.st;
    typedef DeltaStage Stage;
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto pack = [&](size_t u) {                                // A(u): queue the pack kernel + the count copy of unit u
        const Unit &U = units[u]; Stage &S = st[u & 1];
        uint32_t nj = 0, max_groups = 0;
        for (int s = U.s_lo; s < U.s_hi; s++) {
            const StreamRec &R = b->streams[s];
            if (R.n_frames <= U.k) continue;
            const FrameRec &F = b->frames[(size_t)(R.first_frame + U.k)];
            DeltaJob J{};
            J.cur = b->d_out + F.out_off;
            J.prev = U.k > 0 ? b->d_out + b->frames[(size_t)(R.first_frame + U.k - 1)].out_off : nullptr;
            J.X = (uint32_t)R.w; J.Y = (uint32_t)R.h; J.nbx = (uint32_t)((R.w + 15) / 16); J.nb = J.nbx * (uint32_t)((R.h + 15) / 16);
            J.index = nj;
            max_groups = std::max(max_groups, (J.nb + DELTA_WARPS - 1) / DELTA_WARPS);
            S.h_jobs[nj++] = J;
        }
        if (!nj) { *S.h_cnt = 0; return JSP_CUDA(cudaEventRecord(S.packed, b->st_compute)); }
        bool r = JSP_CUDA(cudaMemcpyAsync(S.d_jobs, S.h_jobs, nj * sizeof(DeltaJob), cudaMemcpyHostToDevice, b->st_compute));
        r = r && JSP_CUDA(cudaMemsetAsync(S.d_cnt, 0, 4, b->st_compute));
        delta_pack_kernel<<<dim3(max_groups, nj), DELTA_WARPS * 32, 0, b->st_compute>>>(S.d_jobs, S.d_ent, S.d_pix, S.d_cnt, (uint32_t)worst_unit);
        r = r && JSP_CUDA(cudaGetLastError());
        r = r && JSP_CUDA(cudaMemcpyAsync(S.h_cnt, S.d_cnt, 4, cudaMemcpyDeviceToHost, b->st_compute));
        r = r && JSP_CUDA(cudaEventRecord(S.packed, b->st_compute));
        return r;
    };
    std::vector<uint32_t> counts(units.size(), 0);
    auto fetch = [&](size_t u) {                               // B(u): the count is known -> queue the D2H of exactly that many blocks
        Stage &S = st[u & 1];
        bool r = JSP_CUDA(cudaEventSynchronize(S.packed));
        const uint32_t n = r ? std::min<uint32_t>(*S.h_cnt, (uint32_t)worst_unit) : 0;
        counts[u] = n;
        b->delta_d2h_bytes += 4 + (uint64_t)n * (sizeof(DeltaEntry) + 1024);
        if (r && n) {
            r = r && JSP_CUDA(cudaStreamWaitEvent(b->st_out, S.packed, 0));
            r = r && JSP_CUDA(cudaMemcpyAsync(S.h_ent, S.d_ent, (size_t)n * sizeof(DeltaEntry), cudaMemcpyDeviceToHost, b->st_out));
            r = r && JSP_CUDA(cudaMemcpyAsync(S.h_pix, S.d_pix, (size_t)n * 1024, cudaMemcpyDeviceToHost, b->st_out));
        }
        return r && JSP_CUDA(cudaEventRecord(S.copied, b->st_out));
    };
    // host threads that patch blocks into the stream pictures (started once per call; entries of a unit are unique
    // (stream, block) pairs, so any split of the entry list is race-free)
    struct Pool {
        std::vector<std::thread> th;
        std::mutex mu; std::condition_variable cv_go, cv_done;
        std::function<void(unsigned)> fn; unsigned gen = 0, pending = 0; bool stop = false;
        void start(unsigned n) {
            for (unsigned t = 0; t < n; t++) th.emplace_back([this, t] {
                unsigned seen = 0;
                for (;;) {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_go.wait(lk, [&] { return stop || gen != seen; });
                    if (stop) return;
                    seen = gen;
                    auto f = fn;
                    lk.unlock();
                    f(t + 1);
                    lk.lock();
                    if (--pending == 0) cv_done.notify_one();
                }
            });
        }
        void run(const std::function<void(unsigned)> &f) {      // f(0) on the caller, f(1..n) on the pool
            { std::lock_guard<std::mutex> lk(mu); fn = f; pending = (unsigned)th.size(); gen++; }
            cv_go.notify_all();
            f(0);
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return pending == 0; });
        }
        ~Pool() { { std::lock_guard<std::mutex> lk(mu); stop = true; } cv_go.notify_all(); for (auto &t : th) t.join(); }
    } pool;
    pool.start(hw - 1);
    std::vector<int> jmap;
    auto apply = [&](size_t u) {                               // C(u): the blocks are on the host -> patch them in, report the frames
        const Unit &U = units[u]; Stage &S = st[u & 1];
        if (!JSP_CUDA(cudaEventSynchronize(S.copied))) return false;
        const uint32_t n = counts[u];
        if (n) {
            jmap.clear();                                       // job -> stream
            for (int s = U.s_lo; s < U.s_hi; s++) if (b->streams[s].n_frames > U.k) jmap.push_back(s);
            const unsigned T = n < 2048 ? 1u : hw;
            auto work = [&](unsigned t) {
                if (t >= T) return;
                const uint32_t lo = (uint32_t)((uint64_t)n * t / T), hi = (uint32_t)((uint64_t)n * (t + 1) / T);
                for (uint32_t e = lo; e < hi; e++) {
                    const DeltaEntry E = S.h_ent[e];
                    const StreamRec &R = b->streams[jmap[E.job]];
                    int32_t *pic = stream_pictures[jmap[E.job]];
                    if (!pic) continue;
                    // MSVideo1 never writes the width / height remainder mod 4: those pixels of the caller's buffer stay
                    const int W = R.codec == JSP_CODEC_SCREENPRESSOR ? R.w : (R.w & ~3), H = R.codec == JSP_CODEC_SCREENPRESSOR ? R.h : (R.h & ~3);
                    const int nbx = (R.w + 15) / 16, by = (int)E.block / nbx, bx = (int)E.block - by * nbx;
                    const int cols = std::min(16, W - bx * 16), rows = std::min(16, H - by * 16);
                    if (cols <= 0) continue;
                    const int32_t *src = S.h_pix + (size_t)e * 256;
                    int32_t *dst = pic + (size_t)(by * 16) * R.w + bx * 16;
                    if (cols == 16) for (int r = 0; r < rows; r++) memcpy(dst + (size_t)r * R.w, src + r * 16, 64);
                    else for (int r = 0; r < rows; r++) memcpy(dst + (size_t)r * R.w, src + r * 16, (size_t)cols * 4);
                }
            };
            if (T == 1) work(0); else pool.run(work);
        }
        if (on_frame)                                          // a stream has one unit per frame index: its frame k is complete
            for (int s = U.s_lo; s < U.s_hi; s++)
                if (b->streams[s].n_frames > U.k && stream_pictures[s])
                    on_frame(user, s, U.k, stream_pictures[s], fl[(size_t)(b->streams[s].first_frame + U.k)]);
        return true;
    };
    // software pipeline over the two staging sets: while unit u is patched in on the host, unit u+1 crosses PCIe and -- once
    // u is done with its set -- unit u+2 is packed on the device
    const size_t nu = units.size();
    ok = ok && pack(0) && fetch(0);
    if (ok && nu > 1) ok = pack(1);
    for (size_t u = 0; ok && u < nu; u++) {
        if (u + 1 < nu) ok = ok && fetch(u + 1);
        ok = ok && apply(u);
        if (u + 2 < nu) ok = ok && pack(u + 2);
    }
    cudaStreamSynchronize(b->st_compute); cudaStreamSynchronize(b->st_out);
    return ok ? 0 : -1;
}
