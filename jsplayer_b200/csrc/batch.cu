// batch.cu -- host side of the batch decode path of libjsplayer_cuda (C ABI: include/jsplayer_cuda.h).
//
// Takes the per-stream frame tables the reference's loaders build (src/DataLoader.hx:31 `frames`,
// src/VideoData.hx:68-73 CompressedFrame) for MANY streams, lays bitstreams and output pictures out in
// HBM, orders frames into dependency levels (a frame that copies from the previous picture runs one
// level after it; key frames and frames of other streams run concurrently -- the GOP independence the
// reference relies on when seeking, src/Manager.hx:244-249) and launches the sm_100a kernels.
// No CPU decode path exists here: without a CUDA device every entry point fails.
#include "batch.cuh"
#include <algorithm>
#include <map>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <thread>

namespace jsp {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
bool cuda_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    set_error("CUDA error %s: %s", cudaGetErrorString(e), what);
    return false;
}

// Pinned allocations handed out by jsp_host_alloc: two bitstream ranges may be merged into one copy only when
// they lie in the same allocation (neighbouring allocations can be adjacent in the address space).
static std::mutex g_alloc_mu;
static std::map<uintptr_t, size_t> g_allocs;
static uintptr_t alloc_of(const void *p)
{
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    auto it = g_allocs.upper_bound(a);
    if (it == g_allocs.begin()) return 0;
    --it;
    return a < it->first + it->second ? it->first : 0;
}
static bool same_buffer(const uint8_t *base_a, const uint8_t *pa, const uint8_t *base_b, const uint8_t *pb)
{
    if (base_a == base_b) return true;                   // the caller described both with one buffer
    const uintptr_t x = alloc_of(pa);
    return x != 0 && x == alloc_of(pb);
}

template <typename T>
static bool grow(T *&p, size_t &cap, size_t need, bool zero = false)
{
    if (need <= cap && p) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t n = need + need / 8 + 64;
    if (!JSP_CUDA(cudaMalloc((void **)&p, n * sizeof(T)))) return false;
    if (zero && !JSP_CUDA(cudaMemset(p, 0, n * sizeof(T)))) return false;
    cap = n;
    return true;
}

// MSVideo1.hx:86-104
static bool msv16_just_skips(const uint8_t *src, uint32_t len, uint32_t nblocks)
{
    uint32_t si = 0, n = 0;
    while (si < len) {
        if (si + 1 >= len) return false;                 // (undefined & 0xFC) != 0x84
        const uint32_t a = src[si], b = src[si + 1];
        if ((b & 0xFC) != 0x84) return false;
        n += ((b - 0x84) << 8) + a;
        if (n >= nblocks) return true;
        si += 2;
    }
    return true;
}

// MSVideo1.hx:109-110: the RGB555 decoder returns the previous buffer untouched
bool msv16_unchanged(int w, int h, const uint8_t *src, uint32_t len)
{
    const uint32_t nblocks = (uint32_t)(w >> 2) * (uint32_t)(h >> 2);
    const uint32_t sjs = nblocks / 1023 * 2 + 10;         // MSVideo1.hx:29-30
    return len == 0 || nblocks == 0 || (len < sjs && msv16_just_skips(src, len, nblocks));
}

static int classify(const StreamRec &S, const uint8_t *src, uint32_t len)
{
    if (S.codec == JSP_CODEC_MSVC16) return msv16_unchanged(S.w, S.h, src, len) ? FK_COPY : FK_MSV16;
    if (S.codec == JSP_CODEC_MSVC8) {
        if ((S.w >> 2) == 0 || (S.h >> 2) == 0) return FK_COPY;
        return FK_MSV8;
    }
    return FK_COPY;   // ScreenPressor frames are classified by classify_sp() with the stream's state
}

// What Manager.worker + ScreenPressor.DecompressI/P decide before any entropy decoding happens
// (Manager.hx:505-512, ScreenPressor.hx:129-164, :306-313), replayed on the host for one frame.
static void classify_sp(const StreamRec &S, SpHost &H, FrameRec &R, const uint8_t *src)
{
    R.kind = FK_COPY; R.forced = 0; R.sp_flags = 0; R.fill_value = 0;
    struct SegMark { SpHost &H; FrameRec &R; ~SegMark() {
        // a coded I frame and a model-resetting flat frame start a segment that depends on nothing before it
        R.sp_seg_start = R.kind == FK_SP_I || (R.kind == FK_SP_FLAT && (R.sp_flags & SPJ_RENEW));
        if (R.sp_seg_start) H.n_segments++;
        R.sp_seg = H.n_segments - 1; } } seg_mark{H, R};
    const uint32_t len = R.len;
    if (R.key) {                                              // DecompressI
        if (len == 0) { H.last_flat = false; R.forced = ST_ERROR; return; }   // head undefined -> "unknown version of the codec"
        const int head = src[0], version = (head >> 4) + 1;
        if ((head & 0xF) == 1) {                              // flat (:132-155)
            if (H.version == 0) { R.forced = ST_ERROR; return; }             // ec == null (Appendix E)
            uint32_t c;
            auto rd = [&](uint32_t i) -> uint32_t { return i < len ? src[i] : 0u; };
            if (S.bpp == 16) {
                const uint32_t c16 = len >= 2 ? rd(0) + rd(1) * 256 : 0u;     // src[1] undefined -> NaN -> every `&` gives 0 (:136-141)
                c = (((c16 >> 10) & 0x1F) << 19) | (((c16 >> 5) & 0x1F) << 11) | ((c16 & 0x1F) << 3);
            } else c = len >= 2 ? (rd(3) << 16) | (rd(2) << 8) | rd(1) : 0u;  // `+ b` with b undefined is NaN -> 0; r, g undefined -> 0 under `<<`
            R.kind = FK_SP_FLAT; R.fill_value = c; R.forced = ST_CHANGED;
            R.sp_flags = H.last_flat ? 0u : SPJ_RENEW;        // RenewI skips ec.renewI() after a flat frame (:113)
            H.last_flat = true; H.decodedI = true;
            return;
        }
        H.last_flat = false;
        if ((head & 0xF) != 2) { R.forced = ST_ERROR; return; }              // :157-159
        if (H.version == 0) {
            if (version < 2 || version > 4) { R.forced = ST_ERROR; return; } // initEntro: "unknown version of ScreenPressor!" (:74)
            H.version = version;
        }
        R.kind = FK_SP_I; R.sp_flags = SPJ_IFRAME;
        H.decodedI = true;
    } else {                                                  // DecompressP
        H.last_flat = false;
        if (len == 0 || !H.decodedI || src[0] == 0) return;   // :308-313 -> previous buffer
        R.kind = FK_SP_P;
    }
    if (H.version == 2 && S.bpp == 16) R.sp_flags |= SPJ_DIFF16 | SPJ_CXSHIFT0;   // :59, :200-202 (v3/v4 force shift 2, :71-73)
    if (H.version == 3) R.sp_flags |= SPJ_ANS_V3;
}

struct HostTables {
    std::vector<Msv1Frame> mframes;
    std::vector<Msv1Tile> tile_tab;
    std::vector<CopyJob> jobs;
    std::vector<SpJob> spjobs;
};

// Builds the launch list of `plan` for streams [s_lo, s_hi).
static void build_plan(jsp_batch *b, Plan &plan, int s_lo, int s_hi, HostTables &T, size_t &state_cursor, size_t &ticket_cursor,
                       bool secondary = false)
{
    const uint32_t mframe_base = secondary ? (uint32_t)b->frames.size() : 0u;   // chunk plans use their own frame descriptors
    plan.launches.clear(); plan.finished.clear(); plan.uploads.clear();
    plan.frame_lo = b->streams[s_lo].first_frame;
    plan.frame_hi = b->streams[s_hi - 1].first_frame + b->streams[s_hi - 1].n_frames;
    plan.tile_tab_off = T.tile_tab.size();
    plan.job_off = T.jobs.size();
    plan.spjob_off = T.spjobs.size();
    plan.state_off = state_cursor;
    plan.ticket_off = ticket_cursor;

    int max_level = 0;
    for (int64_t f = plan.frame_lo; f < plan.frame_hi; f++) max_level = std::max(max_level, b->frames[f].level);
    std::vector<std::vector<int64_t>> by_level(max_level + 1);
    for (int64_t f = plan.frame_lo; f < plan.frame_hi; f++) by_level[b->frames[f].level].push_back(f);

    for (int lv = 0; lv <= max_level; lv++) {
        // whole-picture copies (unchanged frames)
        {
            const size_t first = T.jobs.size(); uint32_t maxv = 0;
            std::vector<int64_t> fin;
            for (int64_t f : by_level[lv]) {
                FrameRec &R = b->frames[f];
                const StreamRec &S = b->streams[R.stream];
                // A sparse MSVideo1 inter frame (mostly skip runs: fewer than 3 bytes per block) starts as a wide copy of the
                // previous picture; its decode then only writes the coded blocks.  Inside the decode kernel the skipped blocks
                // are copied by the CTA that owns the bitstream tile -- a 320x240 frame is ONE tile, so one CTA would copy the
                // whole picture (C1: 37 us per frame); many CTAs of the copy kernel do it in a few.
                R.precopy = (R.kind == FK_MSV16 || R.kind == FK_MSV8) && !R.key && R.prev >= 0 && R.level > 0 &&
                            (uint64_t)R.len < 3ull * (uint64_t)(S.w >> 2) * (uint64_t)(S.h >> 2);
                // Fused display store (JSP_BATCH_DISPLAY): the pixels outside the 4x4 block grid of a picture whose size is not
                // a multiple of 4 are written by nobody; the canvas shows them opaque black, so such a picture starts as a fill.
                const bool msv = R.kind == FK_MSV16 || R.kind == FK_MSV8;
                const bool disp = (b->flags & JSP_BATCH_DISPLAY) && S.codec != JSP_CODEC_SCREENPRESSOR;
                const bool edge_fill = disp && msv && !R.precopy && ((S.w | S.h) & 3);
                if (R.kind != FK_COPY && R.kind != FK_SP_P && R.kind != FK_SP_FLAT && !R.precopy && !edge_fill) continue;
                if (S.codec == JSP_CODEC_SCREENPRESSOR) continue;          // ScreenPressor levels are planned below
                if (!R.precopy && !edge_fill) fin.push_back(f);            // a pre-copied picture is final after its decode
                CopyJob J;
                J.dst = b->d_out + R.out_off;
                J.src = R.prev >= 0 ? b->d_out + b->frames[R.prev].out_off : b->ext_prev;
                J.value = disp ? 0xFF000000u : 0u;
                if (edge_fill) J.src = nullptr;
                if (R.kind == FK_SP_FLAT) { J.src = nullptr; J.value = R.fill_value; }
                J.n_vec4 = (uint32_t)(((size_t)S.w * S.h * 4 + 15) / 16);
                maxv = std::max(maxv, J.n_vec4);
                T.jobs.push_back(J);
            }
            if (T.jobs.size() > first) {
                plan.launches.push_back({JSP_K_FRAME_COPY, FK_COPY, first, (uint32_t)(T.jobs.size() - first), maxv, 0});
                plan.finished.push_back(std::move(fin));
            }
        }
        // MSVideo1, one launch per pixel format; tiles are listed tile-major so that a tile's
        // predecessors in its frame were handed out a whole "row" of frames earlier
        for (int kind = FK_MSV16; kind <= FK_MSV8; kind++) {
            std::vector<int64_t> fr;
            uint32_t max_tiles = 0;
            for (int64_t f : by_level[lv])
                if (b->frames[f].kind == kind) { fr.push_back(f); max_tiles = std::max(max_tiles, b->frames[f].n_tiles); }
            if (fr.empty()) continue;
            std::sort(fr.begin(), fr.end(), [&](int64_t x, int64_t y) {
                return b->frames[x].n_tiles != b->frames[y].n_tiles ? b->frames[x].n_tiles > b->frames[y].n_tiles : x < y; });
            const size_t first = T.tile_tab.size();
            for (int64_t f : fr) {
                (secondary ? b->frames[f].state_base2 : b->frames[f].state_base) = (uint32_t)state_cursor;
                state_cursor += b->frames[f].n_tiles;
            }
            size_t live = fr.size();
            for (uint32_t t = 0; t < max_tiles; t++) {
                while (live > 0 && b->frames[fr[live - 1]].n_tiles <= t) live--;
                for (size_t i = 0; i < live; i++) {
                    const FrameRec &R = b->frames[fr[i]];
                    const size_t byte0 = (size_t)t * MSV1_TILE_BYTES;
                    Msv1Tile e;
                    e.src = b->d_bytes + R.d_src + byte0;
                    e.avail = R.len > byte0 ? (uint32_t)std::min<size_t>(MSV1_STAGE_BYTES, R.len - byte0) : 0u;
                    e.frame = (uint32_t)fr[i] + mframe_base;
                    T.tile_tab.push_back(e);
                }
            }
            plan.launches.push_back({JSP_K_MSV1_DECODE, kind, first, (uint32_t)(T.tile_tab.size() - first), 0, (uint32_t)ticket_cursor});
            plan.finished.push_back(fr);
            ticket_cursor++;
        }
    }
    // ScreenPressor: per level, the whole-picture copies / fills that prepare the pictures, then one warp per frame.
    // Range-coder and rANS streams share the launch.  (Measured: dealing the chains -- frames sharing one model-state
    // slot -- into groups with their own CUDA streams, so that levels need not wait for their slowest warp, changes
    // nothing: a batch lasts as long as its slowest CHAIN, 534-545 ms for C4 at 1 to 32 groups.)
    for (int lv = 0; lv <= max_level; lv++) {
        {
            const size_t first = T.jobs.size(); uint32_t maxv = 0;
            std::vector<int64_t> fin;
            for (int64_t f : by_level[lv]) {
                const FrameRec &R = b->frames[f];
                const StreamRec &S = b->streams[R.stream];
                if (S.codec != JSP_CODEC_SCREENPRESSOR) continue;
                if (R.kind != FK_COPY && R.kind != FK_SP_P && R.kind != FK_SP_FLAT) continue;
                if (R.kind != FK_SP_P) fin.push_back(f);                   // a P frame's picture is final after its decode
                CopyJob J;
                J.dst = b->d_out + R.out_off;
                J.src = R.prev >= 0 ? b->d_out + b->frames[R.prev].out_off : b->ext_prev;
                J.value = 0;
                if (R.kind == FK_SP_FLAT) { J.src = nullptr; J.value = R.fill_value; }
                J.n_vec4 = (uint32_t)(((size_t)S.w * S.h * 4 + 15) / 16);
                maxv = std::max(maxv, J.n_vec4);
                T.jobs.push_back(J);
            }
            if (T.jobs.size() > first) {
                plan.launches.push_back({JSP_K_FRAME_COPY, FK_COPY, first, (uint32_t)(T.jobs.size() - first), maxv, 0});
                plan.finished.push_back(std::move(fin));
            }
        }
        const size_t first = T.spjobs.size();
        std::vector<int64_t> fin;
        int n_rc = 0, n_ans = 0; uint32_t max_w = 0;
        // longest frames first: when a launch has more warps than the device holds, the late starters are the short ones
        std::vector<int64_t> order(by_level[lv]);
        std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return b->frames[x].len > b->frames[y].len; });
        // second-generation kernels (sp2_decode.cu): [range-coder I | rANS I | range-coder P and resets | rANS P and resets] -- one
        // launch for the coded I frames, one for the rest, each dealing its two coders to different SMs;
        // first generation (JSP_SP_GEN=1): range-coder frames first
        const bool gen2 = sp_generation() >= 2;
        auto is_rc = [&](int64_t x) { return b->sp_hosts[b->frames[x].stream].version <= 2; };
        auto is_i = [&](int64_t x) { return b->frames[x].kind == FK_SP_I; };
        if (gen2) {
            auto i_end = std::stable_partition(order.begin(), order.end(), is_i);
            std::stable_partition(order.begin(), i_end, is_rc);
            std::stable_partition(i_end, order.end(), is_rc);
        } else {
            std::stable_partition(order.begin(), order.end(), is_rc);
        }
        int n_rc_i = 0, n_ans_i = 0;
        for (int64_t f : order) {
            const FrameRec &R = b->frames[f];
            if (R.kind != FK_SP_I && R.kind != FK_SP_P && !(R.kind == FK_SP_FLAT && (R.sp_flags & SPJ_RENEW))) continue;
            const StreamRec &S = b->streams[R.stream];
            const SpHost &H = b->sp_hosts[R.stream];
            SpJob J{};
            J.src = b->d_bytes + R.d_src; J.len = R.len;
            J.dst = b->d_out + R.out_off;
            J.prev = R.prev >= 0 ? b->d_out + b->frames[R.prev].out_off : b->ext_prev;
            J.status = b->d_status + f;
            J.symbols = b->d_sp_symbols ? b->d_sp_symbols + f : nullptr;
            J.done = b->d_done ? b->d_done + f : nullptr;
            const size_t slot = R.sp_seg > 0 ? (size_t)(R.sp_seg % H.n_slots) : 0;
            J.state = b->d_sp_state + H.state_off + slot * H.state_stride;
            J.bts = b->d_sp_bts + H.bts_off + slot * H.bts_stride;
            J.X = (uint32_t)S.w; J.Y = (uint32_t)S.h; J.flags = R.sp_flags | (H.version > 2 ? SPJ_ANS : 0u);
            J.insign_blocks = (uint32_t)(((S.w + 15) / 16) * ((std::max(0, b->insign_lines) + 15) / 16));
            (H.version > 2 ? n_ans : n_rc)++;
            if (R.kind == FK_SP_I) (H.version <= 2 ? n_rc_i : n_ans_i)++;
            max_w = std::max(max_w, J.X);
            T.spjobs.push_back(J);
            if (R.kind != FK_SP_FLAT) fin.push_back(f);
        }
        if (T.spjobs.size() > first) {
            const int kclass = n_ans == 0 ? JSP_K_SP_ENTROPY_RC : (n_rc == 0 ? JSP_K_SP_ENTROPY_ANS : JSP_K_SP_ENTROPY_MIXED);
            Launch L{kclass, FK_SP_I, first, (uint32_t)(T.spjobs.size() - first), max_w, (uint32_t)ticket_cursor};
            L.n_rc = (uint32_t)n_rc; L.n_rc_i = (uint32_t)n_rc_i; L.n_ans_i = (uint32_t)n_ans_i;
            ticket_cursor += 4;                   // two job queues per launch, two launches (I frames, the rest)
            plan.launches.push_back(L);
            plan.finished.push_back(std::move(fin));
        }
    }
    plan.n_spjobs = T.spjobs.size() - plan.spjob_off;
    plan.n_tile_entries = T.tile_tab.size() - plan.tile_tab_off;
    plan.n_jobs = T.jobs.size() - plan.job_off;
    plan.n_states = state_cursor - plan.state_off;
    plan.n_tickets = ticket_cursor - plan.ticket_off;

    // uploads: one range per stream, merged when the host ranges are (nearly) adjacent
    for (int s = s_lo; s < s_hi; s++) {
        const StreamRec &S = b->streams[s];
        if (S.h_hi <= S.h_lo) continue;
        CopyRange r{S.h_bytes + S.h_lo, S.d_base, (size_t)(S.h_hi - S.h_lo)};
        if (!plan.uploads.empty()) {
            CopyRange &p = plan.uploads.back();
            const uint8_t *pend = p.h + p.bytes;
            if (r.h >= pend && (size_t)(r.h - pend) < 4096 && r.d_off == p.d_off + (size_t)(r.h - p.h) &&
                same_buffer(b->streams[s - 1].h_bytes, pend - 1, S.h_bytes, r.h)) {
                p.bytes = (size_t)(r.h - p.h) + r.bytes;
                continue;
            }
        }
        plan.uploads.push_back(r);
    }
}

// true when the pictures of this stream are stored bottom-up (fused display store with the flip)
static inline bool flipped(const jsp_batch *b, const StreamRec &S)
{
    return (b->flags & JSP_BATCH_DISPLAY) && (b->flags & JSP_BATCH_DISPLAY_FLIP) && S.codec != JSP_CODEC_SCREENPRESSOR;
}

static void fill_mframes(jsp_batch *b, HostTables &T)
{
    const size_t N = b->frames.size();
    T.mframes.assign(b->chunks.empty() ? N : 2 * N, Msv1Frame{});
    for (size_t f = 0; f < N; f++) {
        const FrameRec &R = b->frames[f];
        if (R.kind != FK_MSV16 && R.kind != FK_MSV8) continue;
        const StreamRec &S = b->streams[R.stream];
        Msv1Frame &M = T.mframes[f];
        M.src = b->d_bytes + R.d_src;
        M.out = b->d_out + R.out_off;
        // a frame scheduled at level 0 is not ordered after its predecessor: it must not read it
        M.prev = (R.prev >= 0 && R.level > 0) ? b->d_out + b->frames[R.prev].out_off
                                               : (R.prev < 0 ? b->ext_prev : nullptr);
        M.pal = S.pal_off != SIZE_MAX ? b->d_pal + S.pal_off : nullptr;
        M.status = b->d_status + f;
        M.len = R.len;
        M.X = (uint32_t)S.w;
        M.nbx = (uint32_t)(S.w >> 2);
        M.nblocks = (uint32_t)(S.w >> 2) * (uint32_t)(S.h >> 2);
        M.n_tiles = R.n_tiles;
        M.state_base = R.state_base;
        M.insign_blocks = (uint32_t)std::max(0, (b->insign_lines + 3) >> 2);
        M.flags = (R.prev >= 0 ? MSV1_F_HAS_PRED : 0u) | (R.precopy ? MSV1_F_PRECOPIED : 0u) |
                  ((b->flags & JSP_BATCH_DISPLAY) ? MSV1_F_DISPLAY : 0u) |
                  ((b->flags & JSP_BATCH_DISPLAY) && (b->flags & JSP_BATCH_DISPLAY_FLIP) ? MSV1_F_FLIP : 0u);
        M.Y = (uint32_t)S.h;
        M.inv_nbx = M.nbx > 1 ? (uint32_t)(0x100000000ull / M.nbx) : 0xFFFFFFFFu;
        if (!b->chunks.empty()) { T.mframes[N + f] = M; T.mframes[N + f].state_base = R.state_base2; }
    }
}

static bool upload_tables(jsp_batch *b, HostTables &T, size_t n_states, size_t n_tickets)
{
    if (!grow(b->d_mframes, b->mframes_cap, T.mframes.size())) return false;
    if (!grow(b->d_tile_tab, b->tile_tab_cap, T.tile_tab.size() + 1)) return false;
    if (!grow(b->d_jobs, b->jobs_cap, T.jobs.size() + 1)) return false;
    if (n_states + 1 > b->states_cap) {
        if (b->d_tile_map) cudaFree(b->d_tile_map);
        if (b->d_tile_cnt) cudaFree(b->d_tile_cnt);
        b->d_tile_map = b->d_tile_cnt = nullptr; b->states_cap = 0;
        const size_t n = n_states + n_states / 8 + 64;
        if (!JSP_CUDA(cudaMalloc((void **)&b->d_tile_map, n * 8))) return false;
        if (!JSP_CUDA(cudaMalloc((void **)&b->d_tile_cnt, n * 8))) return false;
        b->states_cap = n;
    }
    if (!grow(b->d_tickets, b->tickets_cap, n_tickets + 1)) return false;
    if (!T.mframes.empty() && !JSP_CUDA(cudaMemcpy(b->d_mframes, T.mframes.data(), T.mframes.size() * sizeof(Msv1Frame), cudaMemcpyHostToDevice))) return false;
    if (!T.tile_tab.empty() && !JSP_CUDA(cudaMemcpy(b->d_tile_tab, T.tile_tab.data(), T.tile_tab.size() * sizeof(Msv1Tile), cudaMemcpyHostToDevice))) return false;
    if (!T.jobs.empty() && !JSP_CUDA(cudaMemcpy(b->d_jobs, T.jobs.data(), T.jobs.size() * sizeof(CopyJob), cudaMemcpyHostToDevice))) return false;
    if (!grow(b->d_spjobs, b->spjobs_cap, T.spjobs.size() + 1)) return false;
    if (!T.spjobs.empty() && !JSP_CUDA(cudaMemcpy(b->d_spjobs, T.spjobs.data(), T.spjobs.size() * sizeof(SpJob), cudaMemcpyHostToDevice))) return false;
    return true;
}

template <class F> static bool run_plan_with(jsp_batch *b, const Plan &P, cudaStream_t st, cudaEvent_t *ev, std::vector<int> *ev_class, F &&after_launch);
static bool run_plan(jsp_batch *b, const Plan &P, cudaStream_t st, cudaEvent_t *ev = nullptr, std::vector<int> *ev_class = nullptr)
{
    return run_plan_with(b, P, st, ev, ev_class, [](int) { return true; });
}

// Enqueues the plan's launches on `st`; after_launch(k) runs on the host right after launch k has been enqueued.
template <class F> static bool run_plan_with(jsp_batch *b, const Plan &P, cudaStream_t st, cudaEvent_t *ev, std::vector<int> *ev_class, F &&after_launch)
{
    if (P.n_states) {
        if (!JSP_CUDA(cudaMemsetAsync(b->d_tile_map + P.state_off, 0, P.n_states * 8, st))) return false;
        if (!JSP_CUDA(cudaMemsetAsync(b->d_tile_cnt + P.state_off, 0, P.n_states * 8, st))) return false;
    }
    if (P.n_tickets && !JSP_CUDA(cudaMemsetAsync(b->d_tickets + P.ticket_off, 0, P.n_tickets * 4, st))) return false;
    if (P.frame_hi > P.frame_lo &&
        !JSP_CUDA(cudaMemsetAsync(b->d_status + P.frame_lo, 0, (size_t)(P.frame_hi - P.frame_lo) * 4, st))) return false;
    if (b->d_sp_symbols && P.frame_hi > P.frame_lo &&
        !JSP_CUDA(cudaMemsetAsync(b->d_sp_symbols + P.frame_lo, 0, (size_t)(P.frame_hi - P.frame_lo) * 4, st))) return false;
    int k = 0;
    for (const Launch &L : P.launches) {
        if (ev) { cudaEventRecord(ev[2 * k], st); ev_class->push_back(L.kclass); }
        switch (L.kclass) {
        case JSP_K_FRAME_COPY:
            launch_frame_copy(b->d_jobs + L.first, L.count, L.max_vec4, b->sm_count, st);
            break;
        case JSP_K_MSV1_DECODE:
            launch_msv1_decode(L.kind == FK_MSV8, (b->flags & JSP_BATCH_DISPLAY) != 0, b->d_mframes, b->d_tile_tab + L.first, L.count,
                               b->d_tile_map, b->d_tile_cnt, b->d_tickets + L.ticket, b->sm_count, st);
            break;
        case JSP_K_SP_ENTROPY_RC: case JSP_K_SP_ENTROPY_ANS: case JSP_K_SP_ENTROPY_MIXED:
            if (sp_generation() < 2)
                launch_sp_decode(b->d_spjobs + L.first, L.count, L.max_vec4, L.n_rc, b->d_tickets + L.ticket, st);
            else if (!launch_sp2_level(b->d_spjobs + L.first, L.n_rc_i, L.n_ans_i, L.n_rc - L.n_rc_i, L.count - L.n_rc - L.n_ans_i,
                                       L.max_vec4, b->d_tickets + L.ticket, st)) {
                set_error("ScreenPressor launch: no side stream on this device"); return false;
            }
            break;
        default: break;
        }
        if (ev) cudaEventRecord(ev[2 * k + 1], st);
        if (!after_launch(k)) return false;
        k++;
    }
    return JSP_CUDA(cudaGetLastError());
}


// the captured replay of jsp_batch_run (see there) holds the addresses of the current tables
static void drop_run_graph(jsp_batch *b)
{
    if (b->run_graph) { cudaGraphExecDestroy(b->run_graph); b->run_graph = nullptr; }
    b->run_graph_failed = false;
}

// (Re)computes dependency levels from the key flags, rebuilds the launch plan and uploads the tables.
static bool plan_and_upload(jsp_batch *b)
{
    drop_run_graph(b);                              // the captured launches hold the old tables' addresses
    for (size_t s = 0; s < b->streams.size(); s++) {
        const StreamRec &S = b->streams[s];
        int level = -1;
        std::vector<int> seg_end;                  // last level of every ScreenPressor segment seen so far
        for (int f = 0; f < S.n_frames; f++) {
            FrameRec &R = b->frames[S.first_frame + f];
            // MSVideo1 key frames do not depend on the previous picture; everything else runs one level later.
            // ScreenPressor: the frames of one segment share model state and run in order; a new segment starts
            // at level 0, or right after the segment whose state slot it reuses.
            bool independent = R.key && (R.kind == FK_MSV16 || R.kind == FK_MSV8);
            int start = 0;
            if (S.codec == JSP_CODEC_SCREENPRESSOR) {
                const SpHost &H = b->sp_hosts[s];
                if (R.sp_seg_start) {
                    independent = true;
                    const int reused = R.sp_seg - H.n_slots;           // the segment that used this state slot before
                    if (reused >= 0 && (size_t)reused < seg_end.size()) start = seg_end[reused] + 1;
                }
                if (R.sp_seg >= 0 && seg_end.size() <= (size_t)R.sp_seg) seg_end.resize(R.sp_seg + 1, 0);
            }
            level = independent ? start : (f == 0 ? 0 : level + 1);
            R.level = level;
            if (S.codec == JSP_CODEC_SCREENPRESSOR && R.sp_seg >= 0) seg_end[R.sp_seg] = level;
        }
    }
    HostTables T;
    size_t state_cursor = 0, ticket_cursor = 0;
    build_plan(b, b->whole, 0, (int)b->streams.size(), T, state_cursor, ticket_cursor);
    // algorithmic bytes per kernel class: a pre-copied sparse MSVideo1 frame is written by the copy kernel (8 B / pixel: read the
    // previous picture, write this one); its decode launch reads the bitstream and rewrites only the coded blocks (not counted)
    memcpy(b->stat_k_bytes, b->stat_k_bytes_base, sizeof b->stat_k_bytes);
    for (const FrameRec &R : b->frames)
        if (R.precopy) {
            const uint64_t npix = (uint64_t)b->streams[R.stream].w * b->streams[R.stream].h;
            b->stat_k_bytes[JSP_K_MSV1_DECODE] -= npix * 4;
            b->stat_k_bytes[JSP_K_FRAME_COPY] += npix * 8;
        }
    // End-to-end path: when the batch is transfer-bound (MSVideo1 only: one wide launch per level, no per-stream
    // serial chains to keep busy) the streams are also cut into chunks of whole streams so that the upload of
    // chunk k+1, the decode of chunk k and the download of chunk k-1 overlap on three CUDA streams.
    b->chunks.clear();
    {
        bool sp = false; size_t out_bytes = 0;
        for (const StreamRec &S : b->streams) { sp = sp || S.codec == JSP_CODEC_SCREENPRESSOR; out_bytes += (size_t)S.w * S.h * 4 * S.n_frames; }
        const size_t target = std::max<size_t>((size_t)96 << 20, out_bytes / 48);
        if (!sp && !b->persist_streams && out_bytes > 4 * target) {
            int s_lo = 0; size_t acc = 0;
            for (int s = 0; s < (int)b->streams.size(); s++) {
                acc += (size_t)b->streams[s].w * b->streams[s].h * 4 * b->streams[s].n_frames;
                if (acc >= target || s + 1 == (int)b->streams.size()) {
                    b->chunks.emplace_back();
                    build_plan(b, b->chunks.back(), s_lo, s + 1, T, state_cursor, ticket_cursor, true);
                    s_lo = s + 1; acc = 0;
                }
            }
        }
    }
    fill_mframes(b, T);
    return upload_tables(b, T, state_cursor, ticket_cursor);
}

// prevFrame bookkeeping + significance (post-pass over the frames of a plan)
static bool run_status(jsp_batch *b, cudaStream_t st)
{
    launch_status_scan(b->d_status, b->d_stream_first, b->d_stream_count, (uint32_t)b->streams.size(), b->ext_has_prev, st);
    const int exact = (b->flags & JSP_BATCH_SIGNIFICANCE) ? 1 : 0;
    if (exact && b->n_sig)
        launch_signif(b->d_sig_cur, b->d_sig_prev, b->d_sig_status, b->d_sig_first, b->d_sig_npx, (uint32_t)b->n_sig, b->sm_count, st);
    if (b->n_kd)
        launch_signif(b->d_kd_cur, b->d_kd_prev, b->d_kd_status, b->d_kd_first, b->d_kd_npx, (uint32_t)b->n_kd, b->sm_count, st, true);
    launch_status_final(b->d_status, b->d_frame_codec, (uint32_t)b->frames.size(), exact, st);
    return JSP_CUDA(cudaGetLastError());
}

}  // namespace jsp

using namespace jsp;

extern "C" {

int jsp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char *jsp_last_error(void) { return g_err; }
const char *jsp_version(void) { return "jsplayer_cuda 0.1 (sm_100a)"; }

void *jsp_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (!JSP_CUDA(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable))) return nullptr;
    { std::lock_guard<std::mutex> lk(g_alloc_mu); g_allocs[reinterpret_cast<uintptr_t>(p)] = bytes ? bytes : 1; }
    return p;
}
void jsp_host_free(void *p)
{
    if (!p) return;
    { std::lock_guard<std::mutex> lk(g_alloc_mu); g_allocs.erase(reinterpret_cast<uintptr_t>(p)); }
    cudaFreeHost(p);
}

jsp_batch *jsp_batch_create(int device, int insignificant_lines, int flags)
{
    int n = jsp_device_count();
    if (n <= 0) { set_error("no CUDA device: libjsplayer_cuda has no CPU fallback"); return nullptr; }
    if (device < 0) { if (!JSP_CUDA(cudaGetDevice(&device))) return nullptr; }
    if (device >= n) { set_error("device %d out of range (%d devices)", device, n); return nullptr; }
    if (!JSP_CUDA(cudaSetDevice(device))) return nullptr;
    // the thread that drives this device, and the pinned buffers it allocates from here on, move to the device's NUMA node
    if (flags & JSP_BATCH_NUMA_BIND) bind_thread_to_device(device);
    jsp_batch *b = new jsp_batch();
    b->device = device; b->insign_lines = insignificant_lines; b->flags = flags;
    cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (!JSP_CUDA(cudaStreamCreateWithFlags(&b->st_compute, cudaStreamNonBlocking)) ||
        !JSP_CUDA(cudaStreamCreateWithFlags(&b->st_in, cudaStreamNonBlocking)) ||
        !JSP_CUDA(cudaStreamCreateWithFlags(&b->st_out, cudaStreamNonBlocking))) { delete b; return nullptr; }
    return b;
}

void jsp_batch_destroy(jsp_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    drop_run_graph(b);
    delta_release(b);
    for (cudaEvent_t e : b->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : b->ev_sync) cudaEventDestroy(e);

    void *ptrs[] = {b->d_bytes, b->d_out, b->d_pal, b->d_status, b->d_mframes, b->d_tile_tab, b->d_jobs, b->d_tile_map,
                    b->d_tile_cnt, b->d_tickets, b->d_sig_cur, b->d_sig_prev, b->d_sig_status, b->d_sig_first, b->d_sig_npx,
                    b->d_stream_first, b->d_stream_count, b->d_frame_codec, b->d_flush, b->d_spjobs, b->d_sp_state,
                    b->d_sp_rows, b->d_sp_bts, b->d_kd_cur, b->d_kd_prev, b->d_kd_status, b->d_kd_first, b->d_kd_npx,
                    b->d_disp, b->d_disp_jobs, b->d_sp_symbols};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (b->h_status) cudaFreeHost(b->h_status);
    if (b->h_done) cudaFreeHost(b->h_done);
    if (b->st_compute) cudaStreamDestroy(b->st_compute);
    if (b->st_in) cudaStreamDestroy(b->st_in);
    if (b->st_out) cudaStreamDestroy(b->st_out);
    delete b;
}

static int64_t batch_configure(jsp_batch *b, const jsp_stream_desc *sd, int n_streams);

// No C++ exception may cross the C ABI: a descriptor table too large for host memory is an error return, not a terminate().
int64_t jsp_batch_configure(jsp_batch *b, const jsp_stream_desc *sd, int n_streams)
{
    try { return batch_configure(b, sd, n_streams); }
    catch (const std::exception &e) { set_error("jsp_batch_configure: %s", e.what()); return -1; }
    catch (...) { set_error("jsp_batch_configure: unknown failure"); return -1; }
}

static int64_t batch_configure(jsp_batch *b, const jsp_stream_desc *sd, int n_streams)
{
    if (!b || !sd || n_streams <= 0) { set_error("jsp_batch_configure: bad arguments"); return -1; }
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    // per-stream codec state survives a re-configure only for the per-stream drop-in (same single stream)
    bool keep = b->persist_streams && (int)b->streams.size() == n_streams;
    for (int s = 0; keep && s < n_streams; s++)
        keep = b->streams[s].codec == sd[s].codec && b->streams[s].w == sd[s].width && b->streams[s].h == sd[s].height &&
               b->streams[s].bpp == sd[s].bpp;
    if (!keep) b->sp_hosts.assign((size_t)n_streams, SpHost{});
    b->streams.clear(); b->frames.clear(); b->chunks.clear();
    size_t bytes_cur = 0, out_cur = 0, pal_cur = 0;
    int64_t nf = 0;
    bool remainder = false;
    b->stat_pixels = b->stat_alg_bytes = b->stat_in_bytes = b->stat_out_bytes = 0;
    memset(b->stat_k_bytes, 0, sizeof b->stat_k_bytes);
    for (int s = 0; s < n_streams; s++) {
        const jsp_stream_desc &D = sd[s];
        if (D.width <= 0 || D.height <= 0 || D.n_frames < 0 || (D.n_frames > 0 && (!D.frame_off || !D.frame_len || !D.frame_key))) {
            set_error("stream %d: bad descriptor", s); return -1;
        }
        // AVI headers carry 32-bit sizes; block counts, tile counts and pixel indices here are sized for real pictures
        if (D.width > 32768 || D.height > 32768 || (int64_t)D.width * D.height > ((int64_t)1 << 28)) {
            set_error("stream %d: picture size %d x %d out of range", s, D.width, D.height); return -1;
        }
        if (D.codec != JSP_CODEC_MSVC16 && D.codec != JSP_CODEC_MSVC8 && D.codec != JSP_CODEC_SCREENPRESSOR) {
            set_error("stream %d: unknown codec %d", s, D.codec); return -1;
        }
        StreamRec S{};
        S.codec = D.codec; S.w = D.width; S.h = D.height; S.bpp = D.bpp; S.n_frames = D.n_frames;
        S.first_frame = nf; S.h_bytes = D.bytes; S.pal_off = SIZE_MAX;
        if (D.codec == JSP_CODEC_MSVC8) { S.pal_off = pal_cur; pal_cur += 256; }
        // a segment cut out of a longer ScreenPressor stream inherits the stream's entropy coder (see jsp_segment_stream)
        if (D.codec == JSP_CODEC_SCREENPRESSOR && !keep && D.sp_version >= 2 && D.sp_version <= 4) b->sp_hosts[s].version = D.sp_version;
        if (D.codec != JSP_CODEC_SCREENPRESSOR && ((D.width & 3) || (D.height & 3))) remainder = true;
        uint64_t lo = UINT64_MAX, hi = 0;
        for (int f = 0; f < D.n_frames; f++) {
            if (D.frame_len[f] == 0) continue;
            lo = std::min<uint64_t>(lo, D.frame_off[f]);
            hi = std::max<uint64_t>(hi, D.frame_off[f] + D.frame_len[f]);
        }
        if (lo == UINT64_MAX) lo = hi = 0;
        if (hi > lo && !D.bytes) { set_error("stream %d: bytes is NULL", s); return -1; }
        S.h_lo = lo; S.h_hi = hi;
        // keep the host address modulo 16 so that 16-byte aligned frames stay aligned in HBM
        const size_t mis = (size_t)(reinterpret_cast<uintptr_t>(D.bytes + lo) & 15u);
        if (s > 0 && b->streams.back().h_hi > b->streams.back().h_lo) {
            // adjacent in host memory => adjacent in HBM (lets uploads merge into large copies)
            const StreamRec &Pv = b->streams.back();
            const uint8_t *pend = Pv.h_bytes + Pv.h_hi;
            const uint8_t *cur = D.bytes + lo;
            if (hi > lo && cur >= pend && (size_t)(cur - pend) < 4096 && same_buffer(Pv.h_bytes, pend - 1, D.bytes, cur)) bytes_cur = Pv.d_base + (size_t)(cur - (Pv.h_bytes + Pv.h_lo));
            else bytes_cur = ((bytes_cur + 255) & ~(size_t)255) + mis;
        } else bytes_cur = ((bytes_cur + 255) & ~(size_t)255) + mis;
        S.d_base = bytes_cur;
        bytes_cur += (size_t)(hi - lo);
        const size_t npix = (size_t)D.width * D.height;
        const size_t npix_pad = (npix + 63) & ~(size_t)63;
        for (int f = 0; f < D.n_frames; f++) {
            FrameRec R{};
            R.stream = s; R.len = D.frame_len[f]; R.key = D.frame_key[f] ? 1 : 0; R.key_in = R.key;
            R.d_src = S.d_base + (size_t)(D.frame_len[f] ? D.frame_off[f] - lo : 0);
            R.out_off = out_cur; out_cur += npix_pad;
            R.prev = f > 0 ? nf + f - 1 : -1;
            const uint8_t *fsrc = D.bytes ? D.bytes + D.frame_off[f] : nullptr;
            R.forced = 0; R.fill_value = 0; R.sp_flags = 0;
            if (S.codec == JSP_CODEC_SCREENPRESSOR) classify_sp(S, b->sp_hosts[s], R, fsrc);
            else R.kind = classify(S, fsrc, R.len);
            R.level = 0;   // assigned by plan_and_upload()
            R.n_tiles = (R.kind == FK_MSV16 || R.kind == FK_MSV8)
                            ? std::max<uint32_t>(1u, (uint32_t)(((size_t)((R.len + 1) >> 1) + MSV1_TILE_WORDS - 1) / MSV1_TILE_WORDS)) : 0u;
            b->frames.push_back(R);
            b->stat_pixels += npix;
            b->stat_alg_bytes += npix * 4 + R.len + ((R.kind == FK_COPY || R.kind == FK_SP_P) ? npix * 4 : 0);
            b->stat_in_bytes += R.len;
            b->stat_out_bytes += npix * 4;
            // per kernel class (DESIGN.md "Algorithmic bytes"): what each kernel must move for this frame
            const int kent = (S.codec == JSP_CODEC_SCREENPRESSOR && b->sp_hosts[s].version > 2) ? JSP_K_SP_ENTROPY_ANS : JSP_K_SP_ENTROPY_RC;
            switch (R.kind) {
            case FK_MSV16: case FK_MSV8: b->stat_k_bytes[JSP_K_MSV1_DECODE] += npix * 4 + R.len; break;
            case FK_COPY:    b->stat_k_bytes[JSP_K_FRAME_COPY] += npix * 8; break;
            case FK_SP_FLAT: b->stat_k_bytes[JSP_K_FRAME_COPY] += npix * 4; break;
            case FK_SP_P:    b->stat_k_bytes[JSP_K_FRAME_COPY] += npix * 8; b->stat_k_bytes[kent] += R.len; break;
            case FK_SP_I:    b->stat_k_bytes[kent] += npix * 4 + R.len; break;
            default: break;
            }
        }
        nf += D.n_frames;
        b->streams.push_back(S);
    }
    if (nf == 0) { set_error("no frames"); return -1; }
    const size_t out_old = b->out_cap;
    if (!grow(b->d_bytes, b->bytes_cap, bytes_cur + 64)) return -1;
    if (!grow(b->d_out, b->out_cap, out_cur, true)) return -1;
    if (remainder && out_old == b->out_cap && !JSP_CUDA(cudaMemset(b->d_out, 0, b->out_cap * 4))) return -1;
    if (!grow(b->d_pal, b->pal_cap, pal_cur + 1)) return -1;
    if (!grow(b->d_status, b->status_cap, (size_t)nf)) return -1;
    b->bytes_used = bytes_cur; b->out_used = out_cur;
    if ((size_t)nf > b->h_status_cap) {
        if (b->h_status) cudaFreeHost(b->h_status);
        b->h_status = nullptr; b->h_status_cap = 0;
        if (!JSP_CUDA(cudaHostAlloc((void **)&b->h_status, (size_t)nf * 4, cudaHostAllocDefault))) return -1;
        b->h_status_cap = (size_t)nf;
    }
    // palettes: MSVideo1.hx:281-291 -- up to 256 little-endian B,G,R,x quads, missing entries 0
    for (int s = 0; s < n_streams; s++) {
        if (b->streams[s].pal_off == SIZE_MAX) continue;
        int32_t pal[256]; memset(pal, 0, sizeof pal);
        const int n = sd[s].palette ? std::min(256, sd[s].palette_bytes / 4) : 0;
        for (int i = 0; i < n; i++) {
            const uint8_t *p = sd[s].palette + 4 * i;
            pal[i] = (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
        }
        if (!JSP_CUDA(cudaMemcpy(b->d_pal + b->streams[s].pal_off, pal, sizeof pal, cudaMemcpyHostToDevice))) return -1;
    }
    // ScreenPressor model state: small tables + 12288 colour rows per stream, block-type scratch
    {
        size_t st_cur = 0, rows_cur = 0, bts_cur = 0; bool any = false;
        size_t n_sp = 0;
        for (int s = 0; s < n_streams; s++) if (b->streams[s].codec == JSP_CODEC_SCREENPRESSOR) n_sp++;
        // at most ~48 GB of model state per device: concurrent segments per stream are capped accordingly
        const int max_slots = (int)std::max<size_t>(1, std::min<size_t>(16, ((size_t)48 << 30) / (std::max<size_t>(1, n_sp) * sp_ans_ctx_bytes())));
        for (int s = 0; s < n_streams; s++) {
            if (b->streams[s].codec != JSP_CODEC_SCREENPRESSOR) continue;
            any = true;
            SpHost &H = b->sp_hosts[s];
            const bool ans = H.version > 2;
            H.n_slots = b->persist_streams ? 1 : std::max(1, std::min(H.n_segments, max_slots));
            H.state_stride = ans ? sp_ans_state_bytes() : (sp_generation() >= 2 ? sp2_rc_state_bytes() : sp_rc_state_bytes());
            H.rows_stride = ((ans ? sp_ans_ctx_bytes() : sp_rc_rows_bytes()) + 255) & ~(size_t)255;
            H.bts_stride = ((size_t)((b->streams[s].w + 15) / 16) * ((b->streams[s].h + 15) / 16) + 255) & ~(size_t)255;
            H.state_off = st_cur; st_cur += H.state_stride * H.n_slots;
            H.rows_off = rows_cur; rows_cur += H.rows_stride * H.n_slots;
            H.bts_off = bts_cur; bts_cur += H.bts_stride * H.n_slots;
        }
        if (any) {
            const bool fresh_alloc = st_cur > b->sp_state_cap || rows_cur > b->sp_rows_cap || !b->d_sp_state;
            if (!grow(b->d_sp_state, b->sp_state_cap, st_cur)) return -1;
            if (!grow(b->d_sp_rows, b->sp_rows_cap, rows_cur)) return -1;
            if (!grow(b->d_sp_bts, b->sp_bts_cap, bts_cur)) return -1;
            if (!grow(b->d_sp_symbols, b->sp_symbols_cap, (size_t)nf, true)) return -1;
            if ((size_t)nf > b->done_cap) {
                if (b->h_done) cudaFreeHost(b->h_done);
                b->h_done = b->d_done = nullptr; b->done_cap = 0;
                if (!JSP_CUDA(cudaHostAlloc((void **)&b->h_done, (size_t)nf * 4, cudaHostAllocMapped))) return -1;
                if (!JSP_CUDA(cudaHostGetDevicePointer((void **)&b->d_done, b->h_done, 0))) return -1;
                b->done_cap = (size_t)nf;
            }
            memset(b->h_done, 0, (size_t)nf * 4);
            if (!keep || fresh_alloc) {
                // generation tags of all rows to 0, generations start at 1: every row reads as "all ones"
                if (!JSP_CUDA(cudaMemsetAsync(b->d_sp_rows, 0, rows_cur, b->st_compute))) return -1;
                if (!JSP_CUDA(cudaMemsetAsync(b->d_sp_state, 0, st_cur, b->st_compute))) return -1;
                for (int s = 0; s < n_streams; s++)
                    if (b->streams[s].codec == JSP_CODEC_SCREENPRESSOR) {
                        const SpHost &H = b->sp_hosts[s];
                        for (int k = 0; k < H.n_slots; k++) {
                            uint8_t *stp = b->d_sp_state + H.state_off + k * H.state_stride, *rwp = b->d_sp_rows + H.rows_off + k * H.rows_stride;
                            if (H.version > 2) sp_ans_state_init(stp, rwp, 1u, b->st_compute);
                            else if (sp_generation() >= 2) sp2_rc_state_init(stp, rwp, 1u, b->st_compute);
                            else sp_rc_state_init(stp, rwp, 1u, b->st_compute);
                        }
                    }
                if (!JSP_CUDA(cudaStreamSynchronize(b->st_compute))) return -1;
            }
        }
    }
    memcpy(b->stat_k_bytes_base, b->stat_k_bytes, sizeof b->stat_k_bytes);
    if (!plan_and_upload(b)) return -1;

    // status post-pass tables
    {
        std::vector<uint32_t> sfirst, scount; std::vector<uint8_t> fcodec(b->frames.size());
        for (const StreamRec &S : b->streams) { sfirst.push_back((uint32_t)S.first_frame); scount.push_back((uint32_t)S.n_frames); }
        std::vector<const int32_t *> cur, prev; std::vector<uint32_t *> stp; std::vector<uint32_t> first, npx;
        for (size_t f = 0; f < b->frames.size(); f++) {
            const FrameRec &R = b->frames[f]; const StreamRec &S = b->streams[R.stream];
            fcodec[f] = (uint8_t)S.codec;
            if (S.codec == JSP_CODEC_MSVC16 && (R.prev >= 0 || b->ext_prev) && R.kind == FK_MSV16) {
                cur.push_back(b->d_out + R.out_off); prev.push_back(R.prev >= 0 ? b->d_out + b->frames[R.prev].out_off : b->ext_prev);
                stp.push_back(b->d_status + f);
                const size_t np = (size_t)S.w * S.h;
                const size_t skip = std::min<size_t>(np, (size_t)std::max(0, b->insign_lines) * S.w);
                // a flipped picture (fused display store) keeps its insignificant lines at the END of the buffer
                if (flipped(b, S)) { first.push_back(0u); npx.push_back((uint32_t)(np - skip)); }
                else               { first.push_back((uint32_t)skip); npx.push_back((uint32_t)np); }
            }
        }
        if (!grow(b->d_stream_first, b->streams_cap, sfirst.size())) return -1;
        if (!grow(b->d_stream_count, b->streams_cap2, scount.size())) return -1;
        if (!grow(b->d_frame_codec, b->frame_codec_cap, fcodec.size())) return -1;
        cudaMemcpy(b->d_stream_first, sfirst.data(), sfirst.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(b->d_stream_count, scount.data(), scount.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(b->d_frame_codec, fcodec.data(), fcodec.size(), cudaMemcpyHostToDevice);
        b->n_sig = cur.size();
        if (b->n_sig > b->sig_cap) {
            void *old[] = {b->d_sig_cur, b->d_sig_prev, b->d_sig_status, b->d_sig_first, b->d_sig_npx};
            for (void *p : old) if (p) cudaFree(p);
            const size_t n = b->n_sig + 64;
            if (!JSP_CUDA(cudaMalloc((void **)&b->d_sig_cur, n * 8)) || !JSP_CUDA(cudaMalloc((void **)&b->d_sig_prev, n * 8)) ||
                !JSP_CUDA(cudaMalloc((void **)&b->d_sig_status, n * 8)) || !JSP_CUDA(cudaMalloc((void **)&b->d_sig_first, n * 4)) ||
                !JSP_CUDA(cudaMalloc((void **)&b->d_sig_npx, n * 4))) return -1;
            b->sig_cap = n;
        }
        if (b->n_sig) {
            cudaMemcpy(b->d_sig_cur, cur.data(), cur.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(b->d_sig_prev, prev.data(), prev.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(b->d_sig_status, stp.data(), stp.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(b->d_sig_first, first.data(), first.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(b->d_sig_npx, npx.data(), npx.size() * 4, cudaMemcpyHostToDevice);
        }
    }
    // key frames: Manager.frames_differ_significantly (Manager.hx:392-421).  Frame 0 of a stream and a key frame that
    // follows a key frame are settled on the host (byte compare of the two compressed frames); a key frame that
    // follows a non-key frame is compared pixel by pixel with the picture before it, on the device.
    {
        std::vector<const int32_t *> cur, prev; std::vector<uint32_t *> stp; std::vector<uint32_t> first, npx;
        for (size_t f = 0; f < b->frames.size(); f++) {
            FrameRec &R = b->frames[f]; const StreamRec &S = b->streams[R.stream];
            if (!R.key) continue;
            const jsp_stream_desc &D = sd[R.stream];
            const int64_t fi = (int64_t)f - S.first_frame;
            if (fi == 0) {
                R.forced |= ST_KEYDIFF;              // `next_frame_to_decode == 0` -> true (Manager.hx:410-411)
            } else if (D.frame_key[fi - 1]) {
                const uint32_t l0 = D.frame_len[fi - 1], l1 = D.frame_len[fi];
                if (l0 != l1 || (l1 && memcmp(D.bytes + D.frame_off[fi - 1], D.bytes + D.frame_off[fi], l1) != 0)) R.forced |= ST_KEYDIFF;
            } else {
                const size_t np = (size_t)S.w * S.h;
                cur.push_back(b->d_out + R.out_off); prev.push_back(b->d_out + b->frames[f - 1].out_off);
                stp.push_back(b->d_status + f);
                const size_t skip = std::min<size_t>(np, (size_t)std::max(0, b->insign_lines) * S.w);
                if (flipped(b, S)) { first.push_back(0u); npx.push_back((uint32_t)(np - skip)); }
                else               { first.push_back((uint32_t)skip); npx.push_back((uint32_t)np); }
            }
        }
        b->n_kd = cur.size();
        if (b->n_kd > b->kd_cap) {
            void *old[] = {b->d_kd_cur, b->d_kd_prev, b->d_kd_status, b->d_kd_first, b->d_kd_npx};
            for (void *p : old) if (p) cudaFree(p);
            const size_t n = b->n_kd + 64;
            if (!JSP_CUDA(cudaMalloc((void **)&b->d_kd_cur, n * 8)) || !JSP_CUDA(cudaMalloc((void **)&b->d_kd_prev, n * 8)) ||
                !JSP_CUDA(cudaMalloc((void **)&b->d_kd_status, n * 8)) || !JSP_CUDA(cudaMalloc((void **)&b->d_kd_first, n * 4)) ||
                !JSP_CUDA(cudaMalloc((void **)&b->d_kd_npx, n * 4))) return -1;
            b->kd_cap = n;
        }
        if (b->n_kd) {
            cudaMemcpy(b->d_kd_cur, cur.data(), cur.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(b->d_kd_prev, prev.data(), prev.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(b->d_kd_status, stp.data(), stp.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(b->d_kd_first, first.data(), first.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(b->d_kd_npx, npx.data(), npx.size() * 4, cudaMemcpyHostToDevice);
        }
    }
    if (!JSP_CUDA(cudaGetLastError())) return -1;
    return nf;
}

int jsp_batch_upload(jsp_batch *b)
{
    if (!b) return -1;
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    for (const CopyRange &r : b->whole.uploads)
        if (!JSP_CUDA(cudaMemcpyAsync(b->d_bytes + r.d_off, r.h, r.bytes, cudaMemcpyHostToDevice, b->st_compute))) return -1;
    return 0;
}

// A plan of hundreds of small dependent launches (ONE MSVideo1 stream: every P frame waits for its predecessor -- BASELINE's
// C1 is 600 launches of a few microseconds each) is bound by the host's launch rate, not by the GPU: such a plan is captured
// into a CUDA graph once and replayed.  Only MSVideo1 / copy launches are captured (the ScreenPressor level launcher creates
// its side streams lazily, which a capture must not see); a plan is re-captured after every re-plan.
static bool graph_worthwhile(const jsp_batch *b)
{
    if (b->whole.launches.size() < 64 || b->persist_streams) return false;
    // ... and only when the launches are small (under 4 Mpixel each on average): replaying big kernels gains nothing
    // (measured: 64 x 1080p streams of 16 frames, 155 launches, 3.5 ms direct / 3.7 ms replayed)
    if (b->stat_pixels / b->whole.launches.size() >= ((uint64_t)4 << 20)) return false;
    for (const Launch &L : b->whole.launches)
        if (L.kclass != JSP_K_MSV1_DECODE && L.kclass != JSP_K_FRAME_COPY) return false;
    return true;
}

int jsp_batch_run(jsp_batch *b)
{
    if (!b) return -1;
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    if (!b->run_graph && !b->run_graph_failed && graph_worthwhile(b)) {
        cudaGraph_t g = nullptr;
        bool ok = cudaStreamBeginCapture(b->st_compute, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            ok = run_plan(b, b->whole, b->st_compute) && run_status(b, b->st_compute);
            ok = (cudaStreamEndCapture(b->st_compute, &g) == cudaSuccess) && ok && g != nullptr;
        }
        ok = ok && cudaGraphInstantiate(&b->run_graph, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        if (!ok) { cudaGetLastError(); b->run_graph = nullptr; b->run_graph_failed = true; }
    }
    if (b->run_graph) return JSP_CUDA(cudaGraphLaunch(b->run_graph, b->st_compute)) ? 0 : -1;
    if (!run_plan(b, b->whole, b->st_compute)) return -1;
    if (!run_status(b, b->st_compute)) return -1;
    return 0;
}

int jsp_batch_sync(jsp_batch *b)
{
    if (!b) return -1;
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    return JSP_CUDA(cudaStreamSynchronize(b->st_compute)) ? 0 : -1;
}

static uint8_t public_flags(uint32_t v)
{
    uint8_t f = 0;
    if (v & ST_CHANGED) f |= JSP_FRAME_CHANGED;
    if (v & ST_SIGNIFICANT) f |= JSP_FRAME_SIGNIFICANT;
    if (v & (ST_ERROR | ST_NEEDS_PREV)) f |= JSP_FRAME_ERROR;
    if (v & ST_KEYDIFF) f |= JSP_FRAME_DIFFERS;
    return f;
}

int jsp_batch_results(jsp_batch *b, uint8_t *flags)
{
    if (!b) return -1;
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    const size_t n = b->frames.size();
    if (!JSP_CUDA(cudaMemcpyAsync(b->h_status, b->d_status, n * 4, cudaMemcpyDeviceToHost, b->st_compute))) return -1;
    if (!JSP_CUDA(cudaStreamSynchronize(b->st_compute))) return -1;
    for (size_t i = 0; i < n; i++) b->h_status[i] |= b->frames[i].forced;
    // A frame the caller flagged as key (so it was scheduled without waiting for its predecessor) that
    // nevertheless copies from the previous picture: demote it, re-level, decode again in order.
    bool demoted = false;
    for (size_t i = 0; i < n; i++)
        if ((b->h_status[i] & ST_NEEDS_PREV) && b->frames[i].key) { b->frames[i].key = 0; demoted = true; }
    if (demoted) {
        if (!plan_and_upload(b)) return -1;
        if (jsp_batch_run(b)) return -1;
        b->rerun_count++;
        return jsp_batch_results(b, flags);
    }
    if (flags) for (size_t i = 0; i < n; i++) flags[i] = public_flags(b->h_status[i]);
    return 0;
}

// Manager.SkipStills (Manager.hx:289-317) asks the loader for the next frame whose change is significant
// (DataLoader.FindPossibleChange, DataLoader.hx:239-252) and decodes forward until it knows.  After a batch decode every
// frame's answer is known: a key frame carries what Manager.worker would have stored for it (frames_differ_significantly,
// JSP_FRAME_DIFFERS), a P frame its PFrameResult.significant_changes.  Returns the first frame >= from_frame of `stream` with
// a significant change, the stream's last frame when there is none (as the reference does), -1 on bad arguments.
int64_t jsp_batch_next_significant(jsp_batch *b, int stream, int64_t from_frame)
{
    if (!b || stream < 0 || stream >= (int)b->streams.size()) { set_error("jsp_batch_next_significant: bad arguments"); return -1; }
    if (jsp_batch_results(b, nullptr)) return -1;
    const StreamRec &S = b->streams[stream];
    if (S.n_frames <= 0) return -1;
    for (int64_t f = from_frame < 0 ? 0 : from_frame; f < S.n_frames; f++) {
        const FrameRec &R = b->frames[(size_t)(S.first_frame + f)];
        const uint8_t fl = public_flags(b->h_status[S.first_frame + f]);
        if (fl & (R.key_in ? JSP_FRAME_DIFFERS : JSP_FRAME_SIGNIFICANT)) return f;
    }
    return S.n_frames - 1;
}

// D2H of pictures [lo, hi) on stream st; adjacent host destinations are merged into one copy
static bool download_range(jsp_batch *b, int64_t lo, int64_t hi, int32_t *const *out_frames, cudaStream_t st)
{
    const int32_t *arena = b->d_out;
    int64_t i = lo;
    while (i < hi) {
        if (!out_frames[i]) { i++; continue; }
        const FrameRec &R = b->frames[i]; const StreamRec &S = b->streams[R.stream];
        const size_t npix = (size_t)S.w * S.h;
        // (a JSP_BATCH_DISPLAY batch delivers such pictures whole: the remainder was filled with the canvas's opaque black)
        const bool rem = S.codec != JSP_CODEC_SCREENPRESSOR && ((S.w & 3) || (S.h & 3)) && !(b->flags & JSP_BATCH_DISPLAY);
        if (rem) {   // the codec never writes the width/height remainder mod 4: leave the caller's pixels alone
            const size_t bw = (size_t)(S.w & ~3), bh = (size_t)(S.h & ~3);
            if (bw && bh && !JSP_CUDA(cudaMemcpy2DAsync(out_frames[i], (size_t)S.w * 4, arena + R.out_off, (size_t)S.w * 4,
                                                        bw * 4, bh, cudaMemcpyDeviceToHost, st))) return false;
            i++; continue;
        }
        int64_t j = i + 1; size_t bytes = npix * 4;
        while (j < hi && out_frames[j] && out_frames[j] == out_frames[j - 1] + ((size_t)b->streams[b->frames[j - 1].stream].w * b->streams[b->frames[j - 1].stream].h) &&
               b->frames[j].out_off == b->frames[j - 1].out_off + ((size_t)b->streams[b->frames[j - 1].stream].w * b->streams[b->frames[j - 1].stream].h) &&
               !((b->streams[b->frames[j].stream].w & 3) || (b->streams[b->frames[j].stream].h & 3)) && bytes < ((size_t)1 << 30)) {
            bytes += (size_t)b->streams[b->frames[j].stream].w * b->streams[b->frames[j].stream].h * 4; j++;
        }
        if (!JSP_CUDA(cudaMemcpyAsync(out_frames[i], arena + R.out_off, bytes, cudaMemcpyDeviceToHost, st))) return false;
        i = j;
    }
    return true;
}

int jsp_batch_download(jsp_batch *b, int32_t *const *out_frames, uint8_t *flags)
{
    if (!b) return -1;
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    if (jsp_batch_results(b, flags)) return -1;
    if (out_frames && !download_range(b, 0, (int64_t)b->frames.size(), out_frames, b->st_compute)) return -1;
    return JSP_CUDA(cudaStreamSynchronize(b->st_compute)) ? 0 : -1;
}

int jsp_batch_download_display(jsp_batch *b, int32_t *const *out_frames, uint8_t *flags, int display_flags)
{
    if (!b || !out_frames) return -1;
    // MSVideo1 pictures of a JSP_BATCH_DISPLAY batch were STORED in the display format by the decode kernel: no pass for them
    const bool fused = (b->flags & JSP_BATCH_DISPLAY) != 0;
    if (fused && ((display_flags & JSP_DISPLAY_FLIP) != 0) != ((b->flags & JSP_BATCH_DISPLAY_FLIP) != 0)) {
        set_error("jsp_batch_download_display: the flip flag differs from the batch's JSP_BATCH_DISPLAY_FLIP"); return -1;
    }
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    if (jsp_batch_results(b, flags)) return -1;
    if (!grow(b->d_disp, b->disp_cap, b->out_used)) return -1;
    std::vector<DisplayJob> jobs; uint32_t maxpx = 0;
    for (size_t i = 0; i < b->frames.size(); i++) {
        if (!out_frames[i]) continue;
        const FrameRec &R = b->frames[i]; const StreamRec &S = b->streams[R.stream];
        if (fused && S.codec != JSP_CODEC_SCREENPRESSOR) continue;
        DisplayJob J;
        J.src = b->d_out + R.out_off; J.dst = b->d_disp + R.out_off; J.X = (uint32_t)S.w; J.Y = (uint32_t)S.h;
        J.from_rgb15 = (S.codec == JSP_CODEC_SCREENPRESSOR && S.bpp == 16) ? 1u : 0u;      // convert_fromRGB15, Manager.hx:120
        J.flip = (display_flags & JSP_DISPLAY_FLIP) ? 1u : 0u;
        maxpx = std::max(maxpx, J.X * J.Y);
        jobs.push_back(J);
    }
    if (!grow(b->d_disp_jobs, b->disp_jobs_cap, jobs.size() + 1)) return -1;
    if (!jobs.empty()) {
        if (!JSP_CUDA(cudaMemcpyAsync(b->d_disp_jobs, jobs.data(), jobs.size() * sizeof(DisplayJob), cudaMemcpyHostToDevice, b->st_compute))) return -1;
        launch_display(b->d_disp_jobs, (uint32_t)jobs.size(), maxpx, b->sm_count, b->st_compute);
        if (!JSP_CUDA(cudaGetLastError())) return -1;
        if (!JSP_CUDA(cudaStreamSynchronize(b->st_compute))) return -1;     // `jobs` is pageable host memory
    }
    // MSVideo1 pictures whose size is not a multiple of 4 are delivered whole here (the remainder shows as black)
    std::vector<int32_t *> outs(out_frames, out_frames + b->frames.size());
    int64_t i = 0; const int64_t n = (int64_t)b->frames.size();
    while (i < n) {
        if (!outs[i]) { i++; continue; }
        const FrameRec &R = b->frames[i]; const StreamRec &S = b->streams[R.stream];
        const int32_t *from = (fused && S.codec != JSP_CODEC_SCREENPRESSOR) ? b->d_out : b->d_disp;
        if (!JSP_CUDA(cudaMemcpyAsync(outs[i], from + R.out_off, (size_t)S.w * S.h * 4, cudaMemcpyDeviceToHost, b->st_compute))) return -1;
        i++;
    }
    return JSP_CUDA(cudaStreamSynchronize(b->st_compute)) ? 0 : -1;
}

// true when the first destination picture is page-locked memory (an asynchronous D2H into pageable memory would
// block the host inside the launch loop)
static bool outputs_pinned(jsp_batch *b, int32_t *const *out_frames)
{
    for (size_t i = 0; i < b->frames.size(); i++) {
        if (!out_frames[i]) continue;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, out_frames[i]) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    }
    return false;
}

// Multi-level batches (inter-frame chains): the pictures a launch finishes are copied out on the download stream
// while the later levels still decode, so the D2H of level L overlaps the decode of levels L+1...
static int decode_host_streamed(jsp_batch *b, int32_t *const *out_frames, uint8_t *flags)
{
    if (jsp_batch_upload(b)) return -1;
    const Plan &P = b->whole;
    while (b->ev_sync.size() < P.launches.size() + 1) { cudaEvent_t e; if (!JSP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming))) return -1; b->ev_sync.push_back(e); }
    cudaEvent_t ev0 = b->ev_sync[P.launches.size()];
    if (!JSP_CUDA(cudaEventRecord(ev0, b->st_compute)) || !JSP_CUDA(cudaStreamWaitEvent(b->st_out, ev0, 0))) return -1;
    std::vector<uint8_t> sent(b->frames.size(), 0);
    const int reruns = b->rerun_count;
    // ScreenPressor pictures are not taken per launch but per FRAME: a launch lasts as long as its longest frame, and the
    // kernel raises a flag in mapped memory when a picture is complete (sp_signal_done) -- the host polls the flags and
    // copies pictures out while the launch is still running
    std::vector<int64_t> polled;
    if (b->h_done) {
        // flags still being raised by an earlier, asynchronous jsp_batch_run must not be mistaken for this run's
        if (!JSP_CUDA(cudaStreamSynchronize(b->st_compute))) return -1;
        memset(b->h_done, 0, b->frames.size() * 4);
    }
    const bool ok = run_plan_with(b, P, b->st_compute, nullptr, nullptr, [&](int k) {
        const std::vector<int64_t> &fin = P.finished[(size_t)k];
        if (fin.empty()) return true;
        const int kc = P.launches[(size_t)k].kclass;
        if (b->h_done && (kc == JSP_K_SP_ENTROPY_RC || kc == JSP_K_SP_ENTROPY_ANS || kc == JSP_K_SP_ENTROPY_MIXED)) {
            polled.insert(polled.end(), fin.begin(), fin.end());
            return true;
        }
        if (!JSP_CUDA(cudaEventRecord(b->ev_sync[(size_t)k], b->st_compute))) return false;
        if (!JSP_CUDA(cudaStreamWaitEvent(b->st_out, b->ev_sync[(size_t)k], 0))) return false;
        for (int64_t f : fin) {
            if (!download_range(b, f, f + 1, out_frames, b->st_out)) return false;
            sent[(size_t)f] = 1;
        }
        return true;
    });
    if (!ok) return -1;
    if (!run_status(b, b->st_compute)) return -1;
    {
        const volatile uint32_t *done = b->h_done;
        size_t left = polled.size();
        while (left > 0) {
            size_t kept = 0;
            for (size_t i = 0; i < left; i++) {
                const int64_t f = polled[i];
                if (done[f]) {
                    if (!download_range(b, f, f + 1, out_frames, b->st_out)) return -1;
                    sent[(size_t)f] = 1;
                } else polled[kept++] = f;
            }
            if (kept == left) {
                // nothing new: stop polling once the compute stream has drained (every flag is up by then -- or the
                // launch failed, which the calls below report)
                if (cudaStreamQuery(b->st_compute) != cudaErrorNotReady) break;
                std::this_thread::yield();
            }
            left = kept;
        }
    }
    if (jsp_batch_results(b, flags)) return -1;                              // syncs the compute stream
    if (b->rerun_count != reruns) std::fill(sent.begin(), sent.end(), 0);    // a demoted key frame: everything was decoded again
    for (size_t f = 0; f < sent.size(); f++)
        if (!sent[f] && !download_range(b, (int64_t)f, (int64_t)f + 1, out_frames, b->st_out)) return -1;
    return JSP_CUDA(cudaStreamSynchronize(b->st_out)) ? 0 : -1;
}

int jsp_batch_decode_host(jsp_batch *b, int32_t *const *out_frames, uint8_t *flags)
{
    if (!b) return -1;
    if (b->chunks.empty() || !out_frames) {
        bool sp_launch = false;                  // a ScreenPressor launch streams its pictures out frame by frame
        for (const Launch &L : b->whole.launches) sp_launch = sp_launch || L.kclass == JSP_K_SP_ENTROPY_RC || L.kclass == JSP_K_SP_ENTROPY_ANS || L.kclass == JSP_K_SP_ENTROPY_MIXED;
        if (out_frames && (b->whole.launches.size() > 1 || (sp_launch && b->h_done)) && outputs_pinned(b, out_frames)) return decode_host_streamed(b, out_frames, flags);
        // single launches and pageable destinations: upload, decode, download back to back
        if (jsp_batch_upload(b)) return -1;
        if (jsp_batch_run(b)) return -1;
        return jsp_batch_download(b, out_frames, flags);
    }
    // transfer-bound batches: three-stream pipeline over chunks of whole streams (PCIe is full duplex)
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    const size_t nc = b->chunks.size();
    const size_t eb = 0;
    while (b->ev_sync.size() < eb + 2 * nc + 2) { cudaEvent_t e; if (!JSP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming))) return -1; b->ev_sync.push_back(e); }
    cudaEvent_t ev0 = b->ev_sync[eb + 2 * nc];
    if (!JSP_CUDA(cudaEventRecord(ev0, b->st_compute))) return -1;           // order after earlier work on the compute stream
    if (!JSP_CUDA(cudaStreamWaitEvent(b->st_in, ev0, 0)) || !JSP_CUDA(cudaStreamWaitEvent(b->st_out, ev0, 0))) return -1;
    const int reruns = b->rerun_count;
    for (size_t c = 0; c < nc; c++) {
        const Plan &P = b->chunks[c];
        for (const CopyRange &r : P.uploads)
            if (!JSP_CUDA(cudaMemcpyAsync(b->d_bytes + r.d_off, r.h, r.bytes, cudaMemcpyHostToDevice, b->st_in))) return -1;
        if (!JSP_CUDA(cudaEventRecord(b->ev_sync[eb + 2 * c], b->st_in))) return -1;
        if (!JSP_CUDA(cudaStreamWaitEvent(b->st_compute, b->ev_sync[eb + 2 * c], 0))) return -1;
        if (!run_plan(b, P, b->st_compute)) return -1;
        if (!JSP_CUDA(cudaEventRecord(b->ev_sync[eb + 2 * c + 1], b->st_compute))) return -1;
        if (!JSP_CUDA(cudaStreamWaitEvent(b->st_out, b->ev_sync[eb + 2 * c + 1], 0))) return -1;
        if (!download_range(b, P.frame_lo, P.frame_hi, out_frames, b->st_out)) return -1;
    }
    if (!run_status(b, b->st_compute)) return -1;
    if (jsp_batch_results(b, flags)) return -1;                              // syncs the compute stream
    if (!JSP_CUDA(cudaStreamSynchronize(b->st_out))) return -1;
    if (b->rerun_count != reruns) {                                           // a mis-flagged key frame was demoted and the batch re-decoded
        if (!download_range(b, 0, (int64_t)b->frames.size(), out_frames, b->st_compute)) return -1;
        return JSP_CUDA(cudaStreamSynchronize(b->st_compute)) ? 0 : -1;
    }
    return 0;
}

uint64_t jsp_batch_device_frame(jsp_batch *b, int64_t i)
{
    if (!b || i < 0 || (size_t)i >= b->frames.size()) return 0;
    return (uint64_t)reinterpret_cast<uintptr_t>(b->d_out + b->frames[i].out_off);
}

int jsp_batch_stats(jsp_batch *b, uint64_t *pixels, uint64_t *alg_bytes, uint64_t *in_bytes, uint64_t *out_bytes)
{
    if (!b) return -1;
    if (pixels) *pixels = b->stat_pixels;
    if (alg_bytes) *alg_bytes = b->stat_alg_bytes;
    if (in_bytes) *in_bytes = b->stat_in_bytes;
    if (out_bytes) *out_bytes = b->stat_out_bytes;
    return 0;
}

int64_t jsp_batch_symbols(jsp_batch *b)
{
    if (!b) return -1;
    if (!b->d_sp_symbols) return 0;
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    std::vector<uint32_t> h(b->frames.size());
    if (!JSP_CUDA(cudaStreamSynchronize(b->st_compute))) return -1;
    if (!JSP_CUDA(cudaMemcpy(h.data(), b->d_sp_symbols, h.size() * 4, cudaMemcpyDeviceToHost))) return -1;
    int64_t n = 0;
    for (uint32_t v : h) n += v;
    return n;
}

int jsp_batch_kernel_bytes(jsp_batch *b, uint64_t *bytes)
{
    if (!b || !bytes) return -1;
    for (int k = 0; k < JSP_N_KERNELS; k++) bytes[k] = b->stat_k_bytes[k];
    if (bytes[JSP_K_SP_ENTROPY_RC] && bytes[JSP_K_SP_ENTROPY_ANS]) {       // both coders present: their jobs share launches
        bytes[JSP_K_SP_ENTROPY_MIXED] = bytes[JSP_K_SP_ENTROPY_RC] + bytes[JSP_K_SP_ENTROPY_ANS];
        bytes[JSP_K_SP_ENTROPY_RC] = bytes[JSP_K_SP_ENTROPY_ANS] = 0;
    }
    return 0;
}

int jsp_batch_time_runs(jsp_batch *b, int warmup, int iters, int flush_l2, float *ms_total, float *kernel_ms, int64_t *launches)
{
    if (!b || iters <= 0) return -1;
    if (!JSP_CUDA(cudaSetDevice(b->device))) return -1;
    cudaStream_t st = b->st_compute;
    if (flush_l2 && !b->d_flush) {
        b->flush_bytes = (size_t)512 << 20;      // > 126 MB L2
        if (!JSP_CUDA(cudaMalloc(&b->d_flush, b->flush_bytes))) return -1;
    }
    for (int i = 0; i < warmup; i++) if (jsp_batch_run(b)) return -1;
    if (!JSP_CUDA(cudaStreamSynchronize(st))) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total = 0.f;
    for (int i = 0; i < iters; i++) {
        if (flush_l2) cudaMemsetAsync(b->d_flush, i & 0xFF, b->flush_bytes, st);
        cudaEventRecord(e0, st);
        if (jsp_batch_run(b)) return -1;
        cudaEventRecord(e1, st);
        if (!JSP_CUDA(cudaEventSynchronize(e1))) return -1;
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        total += ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_total) *ms_total = total;
    if (kernel_ms || launches) {
        // attribution pass: one extra run with an event in front of every launch
        const size_t n = b->whole.launches.size();
        while (b->ev_pool.size() < 2 * n + 2) { cudaEvent_t e; cudaEventCreate(&e); b->ev_pool.push_back(e); }
        float acc[JSP_N_KERNELS] = {0}; int64_t cnt[JSP_N_KERNELS] = {0};
        for (int i = 0; i < iters; i++) {
            if (flush_l2) cudaMemsetAsync(b->d_flush, i & 0xFF, b->flush_bytes, st);
            std::vector<int> cls;
            if (!run_plan(b, b->whole, st, b->ev_pool.data(), &cls)) return -1;
            if (!JSP_CUDA(cudaStreamSynchronize(st))) return -1;
            for (size_t k = 0; k < cls.size(); k++) {
                float ms = 0.f; cudaEventElapsedTime(&ms, b->ev_pool[2 * k], b->ev_pool[2 * k + 1]);
                acc[cls[k]] += ms; cnt[cls[k]]++;
            }
            if (!run_status(b, st)) return -1;
        }
        for (int k = 0; k < JSP_N_KERNELS; k++) { if (kernel_ms) kernel_ms[k] = acc[k]; if (launches) launches[k] = cnt[k]; }
    }
    return 0;
}

int jsp_batch_decode(const jsp_stream_desc *streams, int n_streams, int n_gpus, int32_t *const *out_frames,
                     uint8_t *out_changed, uint8_t *out_significant, int32_t *out_status)
{
    if (!streams || n_streams <= 0) { set_error("jsp_batch_decode: bad arguments"); return -1; }
    const int ndev = jsp_device_count();
    if (ndev <= 0) { set_error("no CUDA device: libjsplayer_cuda has no CPU fallback"); return -1; }
    if (n_gpus <= 0) n_gpus = 1;
    n_gpus = std::min(std::min(n_gpus, ndev), n_streams);
    // longest-first round robin of whole streams over the devices: no exchange step exists
    std::vector<int64_t> first(n_streams + 1, 0);
    std::vector<uint64_t> weight(n_streams, 0);
    for (int s = 0; s < n_streams; s++) {
        first[s + 1] = first[s] + streams[s].n_frames;
        for (int f = 0; f < streams[s].n_frames; f++) weight[s] += streams[s].frame_len[f] + (uint64_t)streams[s].width * streams[s].height / 8;
    }
    std::vector<int> order(n_streams);
    for (int s = 0; s < n_streams; s++) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return weight[x] > weight[y]; });
    std::vector<std::vector<int>> shard(n_gpus);
    std::vector<uint64_t> load(n_gpus, 0);
    for (int s : order) {
        int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        shard[g].push_back(s); load[g] += weight[s];
    }
    std::vector<int> rc(n_gpus, 0);
    std::vector<std::string> errs(n_gpus);
    auto work = [&](int g) {
      try {
        std::sort(shard[g].begin(), shard[g].end());
        std::vector<jsp_stream_desc> sd; std::vector<int32_t *> outs; std::vector<int64_t> gidx;
        for (int s : shard[g]) {
            sd.push_back(streams[s]);
            for (int f = 0; f < streams[s].n_frames; f++) { outs.push_back(out_frames ? out_frames[first[s] + f] : nullptr); gidx.push_back(first[s] + f); }
        }
        if (sd.empty()) return;
        // with several GPUs every device has its own host thread (created below), which may be moved to the device's NUMA node
        jsp_batch *b = jsp_batch_create(g, 0, JSP_BATCH_SIGNIFICANCE | (n_gpus > 1 ? JSP_BATCH_NUMA_BIND : 0));
        std::vector<uint8_t> fl(outs.size());
        if (!b || jsp_batch_configure(b, sd.data(), (int)sd.size()) < 0 || jsp_batch_decode_host(b, outs.data(), fl.data())) {
            rc[g] = -1; errs[g] = jsp_last_error();
        } else {
            for (size_t i = 0; i < gidx.size(); i++) {
                if (out_changed) out_changed[gidx[i]] = (fl[i] & JSP_FRAME_CHANGED) ? 1 : 0;
                if (out_significant) out_significant[gidx[i]] = (fl[i] & JSP_FRAME_SIGNIFICANT) ? 1 : 0;
                if (out_status) out_status[gidx[i]] = (fl[i] & JSP_FRAME_ERROR) ? JSP_ERROR_OCCURED : JSP_ZERO_STATE;
            }
        }
        jsp_batch_destroy(b);
      } catch (const std::exception &e) { rc[g] = -1; errs[g] = e.what(); }
        catch (...) { rc[g] = -1; errs[g] = "unknown failure"; }
    };
    std::vector<std::thread> th;
    if (n_gpus == 1) work(0);                              // on the caller's thread, whose CPU affinity is left alone
    else for (int g = 0; g < n_gpus; g++) th.emplace_back(work, g);
    for (auto &t : th) t.join();
    for (int g = 0; g < n_gpus; g++) if (rc[g]) { set_error("gpu %d: %s", g, errs[g].c_str()); return -1; }
    return 0;
}

}  // extern "C"
