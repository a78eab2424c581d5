// batch.cuh -- host-side batcher of libjsplayer_cuda: frame / GOP tables, launch plans, device arenas.
// Plays the role DataLoader's frame table + Manager.worker's frame loop play in the reference
// (src/DataLoader.hx:31,93-98; src/Manager.hx:454-525), for many streams at once.
#pragma once
#include "common.cuh"
#include "sp_common.cuh"
#include "../../include/jsplayer_cuda.h"
#include <string>
#include <vector>

struct jsp_batch;
namespace jsp {

void set_error(const char *fmt, ...);
bool cuda_ok(cudaError_t e, const char *what);
bool msv16_unchanged(int w, int h, const uint8_t *src, uint32_t len);
#define JSP_CUDA(call) ::jsp::cuda_ok((call), #call)

enum FrameKind : int {
    FK_MSV16 = 0, FK_MSV8 = 1,
    FK_COPY = 2,        // picture = previous picture (unchanged / skipped / failed frame)
    FK_SP_I = 3,        // ScreenPressor coded I frame
    FK_SP_P = 4,        // ScreenPressor P frame (starts from a copy of the previous picture)
    FK_SP_FLAT = 5,     // ScreenPressor flat I frame: whole-picture fill (+ model reset)
};

// ScreenPressor per-stream bookkeeping the reference keeps in the codec object (ScreenPressor.hx:26-43)
struct SpHost {
    int version = 0;            // 0 = no entropy coder yet (`ec == null`), else 2/3/4 (initEntro, :66-79)
    bool decodedI = false;
    bool last_flat = false;     // last_one_was_flat != null
    // model state slots: every independent segment (coded I frame / model-resetting flat frame up to the next
    // one) decodes with its own state, so the segments of one stream run concurrently; slot = segment % n_slots
    size_t state_off = 0, rows_off = 0, bts_off = 0;
    size_t state_stride = 0, rows_stride = 0, bts_stride = 0;
    int n_slots = 1, n_segments = 0;
};

// sp_decode.cu (sp_rc.cuh)
size_t sp_rc_state_bytes();
size_t sp_rc_rows_bytes();
void sp_rc_state_init(void *d_state, void *d_rows, uint32_t gen0, cudaStream_t st);
// sp_decode.cu (sp_ans.cuh)
size_t sp_ans_state_bytes();
size_t sp_ans_ctx_bytes();
void sp_ans_state_init(void *d_state, void *d_ctx, uint32_t gen0, cudaStream_t st);
void launch_sp_decode(const SpJob *d_jobs, uint32_t n_jobs, uint32_t max_width, uint32_t n_rc, uint32_t *d_queue, cudaStream_t st);   // sp_decode.cu: both coders, one launch
// sp2_decode.cu: second-generation kernels (JSP_SP_GEN=1 in the environment selects the first generation for A/B runs)
int bind_thread_to_device(int device);                     // numa_bind.cpp
int sp_generation();
size_t sp2_rc_state_bytes();
void sp2_rc_state_init(void *d_state, void *d_rows, uint32_t gen0, cudaStream_t st);
bool launch_sp2_level(const SpJob *d_jobs, uint32_t n_rc_i, uint32_t n_ans_i, uint32_t n_rc_p, uint32_t n_ans_p, uint32_t max_width,
                      uint32_t *d_queue, cudaStream_t st);

struct DeltaStages;
struct StreamRec {
    int codec, w, h, bpp;
    int n_frames;
    int64_t first_frame;            // global index of frame 0
    const uint8_t *h_bytes;         // host base pointer
    uint64_t h_lo, h_hi;            // byte range [lo, hi) of h_bytes that holds frames
    size_t d_base;                  // device offset of h_bytes + h_lo
    size_t pal_off;                 // palette slot (int32 index into d_pal) or SIZE_MAX
};

struct FrameRec {
    int stream;
    uint32_t len;
    size_t d_src;                   // offset into d_bytes
    size_t out_off;                 // int32 offset into d_out
    int64_t prev;                   // global index of the previous frame of the stream, -1 if none
    int level;
    int kind;
    uint8_t key;
    bool precopy = false;           // sparse MSVideo1 inter frame: the level's copy launch writes the previous picture first
    uint8_t key_in;                 // the caller's key flag (key may be demoted: a "key" frame that copies from its predecessor)
    uint32_t n_tiles, state_base;
    uint32_t state_base2;           // the frame's tile-state slice in the chunked (pipelined) plans
    uint32_t forced;                // status bits decided on the host (ST_ERROR, ST_CHANGED of flat frames)
    uint32_t fill_value;            // FK_SP_FLAT colour
    uint32_t sp_flags;              // SPJ_* for ScreenPressor jobs
    bool sp_seg_start;              // ScreenPressor: the frame starts an independent segment
    int sp_seg;                     // ScreenPressor: index of the independent segment the frame belongs to (-1: none yet)
};

struct Launch {
    int kclass;                     // JSP_K_*
    int kind;                       // FrameKind
    size_t first;                   // offset into the plan's tile table / job table
    uint32_t count;                 // CTAs / jobs
    uint32_t max_vec4;              // copy jobs: largest job
    uint32_t ticket;                // ticket counter slot (MSVideo1 tiles; two job queues of a mixed ScreenPressor launch)
    uint32_t n_rc = 0;              // ScreenPressor: the first n_rc jobs are range-coder frames
    uint32_t n_rc_i = 0, n_ans_i = 0;   // ... and each coder's jobs start with its coded I frames (second-generation kernels: sp2_decode.cu)
};

struct CopyRange { const uint8_t *h; size_t d_off; size_t bytes; };

// One schedulable unit: a set of streams with everything needed to upload, decode and download them.
struct Plan {
    int64_t frame_lo = 0, frame_hi = 0;        // global frame range covered (streams are contiguous)
    std::vector<Launch> launches;
    std::vector<std::vector<int64_t>> finished;   // per launch: frames whose picture is final once it completes
    std::vector<CopyRange> uploads;
    size_t tile_tab_off = 0, n_tile_entries = 0;   // slice of d_tile_tab
    size_t job_off = 0, n_jobs = 0;                // slice of d_jobs
    size_t state_off = 0, n_states = 0;            // slice of the tile-state arrays
    size_t ticket_off = 0, n_tickets = 0;
    size_t spjob_off = 0, n_spjobs = 0;            // slice of d_spjobs
};

void delta_release(struct ::jsp_batch *b);                  // delta.cu
}  // namespace jsp

struct jsp_batch {
    int device = 0;
    int sm_count = 148;
    int insign_lines = 0;
    int flags = 0;
    cudaStream_t st_compute = nullptr, st_in = nullptr, st_out = nullptr;
    std::vector<cudaEvent_t> ev_pool;              // timing events (jsp_batch_time_runs)
    std::vector<cudaEvent_t> ev_sync;              // ordering-only events (end-to-end pipeline)


    std::vector<jsp::StreamRec> streams;
    std::vector<jsp::FrameRec> frames;
    jsp::Plan whole;
    std::vector<jsp::Plan> chunks;

    // device arenas (grown on demand, never shrunk)
    uint8_t *d_bytes = nullptr;   size_t bytes_cap = 0, bytes_used = 0;
    int32_t *d_out = nullptr;     size_t out_cap = 0, out_used = 0;      // in int32
    int32_t *d_pal = nullptr;     size_t pal_cap = 0;
    uint32_t *d_status = nullptr; size_t status_cap = 0;
    jsp::Msv1Frame *d_mframes = nullptr; size_t mframes_cap = 0;
    jsp::Msv1Tile *d_tile_tab = nullptr;  size_t tile_tab_cap = 0;
    jsp::CopyJob *d_jobs = nullptr; size_t jobs_cap = 0;
    unsigned long long *d_tile_map = nullptr, *d_tile_cnt = nullptr; size_t states_cap = 0;
    unsigned int *d_tickets = nullptr; size_t tickets_cap = 0;
    // ScreenPressor: per-stream model state in HBM
    std::vector<jsp::SpHost> sp_hosts;             // one per stream (unused entries for MSVideo1 streams)
    jsp::SpJob *d_spjobs = nullptr; size_t spjobs_cap = 0;
    uint8_t *d_sp_state = nullptr; size_t sp_state_cap = 0;
    uint8_t *d_sp_rows = nullptr;  size_t sp_rows_cap = 0;
    uint8_t *d_sp_bts = nullptr;   size_t sp_bts_cap = 0;
    uint32_t *d_sp_symbols = nullptr; size_t sp_symbols_cap = 0;   // per frame: entropy-coded symbols decoded (reporting)
    int persist_streams = 0;                       // per-stream drop-in: keep codec state across configure calls
    // significance post-pass tables
    const int32_t **d_sig_cur = nullptr; const int32_t **d_sig_prev = nullptr; uint32_t **d_sig_status = nullptr;
    uint32_t *d_sig_first = nullptr, *d_sig_npx = nullptr; size_t sig_cap = 0, n_sig = 0;
    uint32_t *d_stream_first = nullptr, *d_stream_count = nullptr; size_t streams_cap = 0, streams_cap2 = 0;
    uint8_t *d_frame_codec = nullptr; size_t frame_codec_cap = 0;
    void *d_flush = nullptr; size_t flush_bytes = 0;
    // key-frame change detection (Manager.frames_differ_significantly): pixel-compare jobs
    const int32_t **d_kd_cur = nullptr; const int32_t **d_kd_prev = nullptr; uint32_t **d_kd_status = nullptr;
    uint32_t *d_kd_first = nullptr, *d_kd_npx = nullptr; size_t kd_cap = 0, n_kd = 0;
    // display epilogue
    int32_t *d_disp = nullptr; size_t disp_cap = 0;
    jsp::DisplayJob *d_disp_jobs = nullptr; size_t disp_jobs_cap = 0;

    uint32_t *h_status = nullptr; size_t h_status_cap = 0;   // pinned
    uint32_t *h_done = nullptr, *d_done = nullptr; size_t done_cap = 0;   // mapped pinned: per-frame completion flags of ScreenPressor frames

    cudaGraphExec_t run_graph = nullptr; // jsp_batch_run of a launch-bound plan (hundreds of small dependent launches), captured once
    bool run_graph_failed = false;
    jsp::DeltaStages *delta = nullptr;   // staging of jsp_batch_decode_host_delta (delta.cu)
    uint64_t delta_d2h_bytes = 0;        // bytes the last jsp_batch_decode_host_delta moved device -> host

    const int32_t *ext_prev = nullptr;   // previous picture held outside the batch (per-stream drop-in)
    int ext_has_prev = 0;                // codec's prevFrame was non-null before the batch's first frame
    int rerun_count = 0;

    uint64_t stat_pixels = 0, stat_alg_bytes = 0, stat_in_bytes = 0, stat_out_bytes = 0;
    uint64_t stat_k_bytes[JSP_N_KERNELS] = {0};    // algorithmic bytes per run, per kernel class (as planned)
    uint64_t stat_k_bytes_base[JSP_N_KERNELS] = {0};   // ... before the planner moved pre-copied frames to the copy kernel
};
