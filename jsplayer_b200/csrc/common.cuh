// common.cuh -- device-side frame descriptors and small helpers shared by the kernels
// and the host batcher of libjsplayer_cuda (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace jsp {

// ---- per-frame result bits accumulated on the device (one u32 per output frame) ----
enum : uint32_t {
    ST_CHANGED      = 1u << 0,   // a coded (non-skip) block was decoded: `changes` (MSVideo1.hx:169,180) / SP frame altered
    ST_SIGNIF_ROWS  = 1u << 1,   // a coded block lies in a block row >= insignificant_blocks (MSVideo1.hx:188-194)
    ST_PIXDIFF      = 1u << 2,   // some pixel from insign_lines*X on differs from the previous picture (:195-204)
    ST_ERROR        = 1u << 3,   // malformed bitstream / DecoderState.error_occured
    ST_NEEDS_PREV   = 1u << 4,   // a frame scheduled as a key frame copied from a previous picture it was not ordered after
    ST_SIGNIFICANT  = 1u << 5,   // final PFrameResult.significant_changes
    ST_HAS_PREV     = 1u << 6,   // codec's prevFrame was non-null when this frame was decoded
    ST_KEYDIFF      = 1u << 7,   // key frame differs significantly from the picture before it (Manager.hx:392-421)
};

// ---- MSVideo1 ----
enum : uint32_t {
    MSV1_F_HAS_PRED = 1u << 0,   // the stream has an earlier frame (prev == nullptr then means "not ordered", not "none")
    MSV1_F_PRECOPIED = 1u << 1,  // the output already holds the previous picture: skip runs need no copy
    MSV1_F_DISPLAY = 1u << 2,    // fused display epilogue (JSP_BATCH_DISPLAY): pixels are stored as canvas bytes R,G,B,A
    MSV1_F_FLIP = 1u << 3,       // ... and picture row y lands in row Y-1-y (JSP_BATCH_DISPLAY_FLIP)
};

struct Msv1Frame {
    const uint8_t *src;      // first compressed byte (device)
    int32_t       *out;      // output picture (device), X*Y int32, bitstream row order
    const int32_t *prev;     // previous picture or nullptr (=> zeros)
    const int32_t *pal;      // 256-entry palette (8-bit codec) or nullptr
    uint32_t      *status;   // result bits of this frame
    uint32_t len;            // compressed bytes
    uint32_t X;              // width = row stride in pixels
    uint32_t nbx;            // X >> 2
    uint32_t nblocks;        // (X>>2)*(Y>>2)
    uint32_t n_tiles;        // bitstream tiles of this frame (>= 1)
    uint32_t state_base;     // first slot of this frame in the tile-state arrays
    uint32_t insign_blocks;  // (insignificant_lines+3)>>2
    uint32_t flags;          // MSV1_F_*
    uint32_t inv_nbx;        // floor(2^32 / nbx): block row = umulhi(block, inv_nbx) (+1 fix-up)
    uint32_t Y;              // picture height (row flip of the fused display store); sizeof == 80: copied in 16-byte units
};

// one 4 KiB bitstream tile of a frame, in launch (ticket) order: everything the prefetch needs in ONE 16-byte load
struct Msv1Tile {
    const uint8_t *src;      // first byte of the tile (device)
    uint32_t avail;          // bytes of the frame from there on, capped at MSV1_STAGE_BYTES
    uint32_t frame;          // index into the frame descriptor table
};

// whole-picture copy / fill jobs (unchanged frames, flat frames, P-frame pre-copies)
struct CopyJob {
    int32_t       *dst;
    const int32_t *src;      // nullptr => fill with `value`
    uint32_t       value;
    uint32_t       n_vec4;   // number of 16-byte units (pictures are padded to 16 B)
};

constexpr int MSV1_THREADS    = 128;
constexpr int MSV1_SEG_WORDS  = 16;                                   // 16-bit words per lane segment (>= 9 = longest opcode)
constexpr int MSV1_TILE_WORDS = MSV1_THREADS * MSV1_SEG_WORDS;        // 2048
constexpr int MSV1_TILE_BYTES = MSV1_TILE_WORDS * 2;                  // 4096
constexpr int MSV1_STAGE_BYTES = MSV1_TILE_BYTES + 32;                // + look-ahead for an opcode that starts in the last word

// host-callable launchers (implemented in the .cu files)
void launch_msv1_decode(bool is8, bool display, const Msv1Frame *d_frames, const Msv1Tile *d_tiles, uint32_t n_tiles,
                        unsigned long long *d_tile_map, unsigned long long *d_tile_cnt,
                        unsigned int *d_ticket, int sm_count, cudaStream_t st);
void launch_frame_copy(const CopyJob *d_jobs, uint32_t n_jobs, uint32_t max_vec4, int sm_count, cudaStream_t st);
void launch_signif(const int32_t *const *d_cur, const int32_t *const *d_prev, uint32_t *const *d_status,
                   const uint32_t *d_first_px, const uint32_t *d_npx, uint32_t n_jobs, int sm_count, cudaStream_t st,
                   bool key_frames = false);

// display epilogue (Manager.fill_bitmap_data, Manager.hx:363-381 + the render-time flip of Main.hx:946)
struct DisplayJob {
    const int32_t *src;
    int32_t       *dst;
    uint32_t X, Y;
    uint32_t from_rgb15;     // ScreenPressor at 16 bpp: 0xFF000000 | (c << 3) instead of the R/B swap
    uint32_t flip;
};
void launch_display(const DisplayJob *d_jobs, uint32_t n_jobs, uint32_t max_pixels, int sm_count, cudaStream_t st);

void launch_status_scan(uint32_t *d_status, const uint32_t *d_stream_first, const uint32_t *d_stream_count,
                        uint32_t n_streams, int init_has_prev, cudaStream_t st);
void launch_status_final(uint32_t *d_status, const uint8_t *d_frame_codec, uint32_t n, int exact, cudaStream_t st);

}  // namespace jsp
