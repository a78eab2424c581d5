/*
 * jsplayer_cuda.h -- C ABI of libjsplayer_cuda, the B200 (sm_100a) batch video
 * decoder that drops in for thedeemon/jsplayer's two codec hot paths.
 *
 * The reference has no FFI (it is Haxe compiled to JavaScript); the boundary
 * replaced here is `interface IVideoCodec` (reference src/IVideoCodec.hx:16-29)
 * as constructed by Manager.video_info_cb (src/Manager.hx:105-111) and driven by
 * Manager.worker (src/Manager.hx:454-525).  Every entry point below cites the
 * member it replaces.  Plain pointers and sizes only; no CUDA or torch types.
 * A Haxe/hxcpp extern for these symbols is in haxe/JsplayerCuda.hx and the
 * reference-side change is described in INTEGRATION.md.
 *
 * There is NO CPU fallback: every decode entry point fails (JSP_ERROR_OCCURED /
 * negative return) when no CUDA device is usable.
 */
#ifndef JSPLAYER_CUDA_H
#define JSPLAYER_CUDA_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define JSP_API __attribute__((visibility("default")))

/* enum DecoderState -- IVideoCodec.hx:5-9 */
typedef enum { JSP_ZERO_STATE = 0, JSP_IN_PROGRESS = 1, JSP_ERROR_OCCURED = 2 } jsp_state;
/* enum CodecType -- VideoData.hx:75-80 */
typedef enum { JSP_CODEC_SCREENPRESSOR = 0, JSP_CODEC_MSVC16 = 1, JSP_CODEC_MSVC8 = 2 } jsp_codec;
/* typedef PFrameResult -- IVideoCodec.hx:11-14. data_pnt is dst, the retained previous
 * buffer (unchanged frame), or NULL (nothing decoded yet). */
typedef struct { int32_t *data_pnt; int32_t significant_changes; } jsp_pframe_result;

typedef struct jsp_dec jsp_dec;

/* ---- library ---- */
JSP_API int         jsp_device_count(void);            /* usable CUDA devices, 0 if none */
JSP_API const char *jsp_last_error(void);              /* thread-local, never NULL */
JSP_API const char *jsp_version(void);
/* pinned host memory for frame / bitstream buffers (what DataLoaderAVIIndexed's frame table holds) */
JSP_API void *jsp_host_alloc(size_t bytes);
JSP_API void  jsp_host_free(void *p);
/* host topology helpers behind JSP_BATCH_NUMA_BIND: the NUMA node of a device (-1 = unknown or single node) and the bind itself
 * for threads the caller owns (returns the node, -1 = nothing changed) */
JSP_API int   jsp_numa_node_of_device(int device);
JSP_API int   jsp_numa_bind_thread(int device);
/* `reps` device -> pinned-host copies of `bytes` each from `device`: GB/s (0 on failure).  Called by N ranks at the same time it
 * measures what this box's host takes from N GPUs at once -- the ceiling of the end-to-end path (4 bytes per decoded pixel). */
JSP_API double jsp_host_d2h_gbs(int device, size_t bytes, int reps);

/* ---- per-stream drop-in: one stateful codec per stream, frames in order, host buffers ----
 * new MSVideo1_16bit(w,h) MSVideo1.hx:20-31 | new MSVideo1_8bit(w,h,palette) :267-274 |
 * new ScreenPressor(w,h,bpp) ScreenPressor.hx:53-64.  palette = strf bytes from offset 40
 * (AVIParser.hx:79-85), little-endian B,G,R,x quads.  device < 0 selects the current device. */
JSP_API jsp_dec  *jsp_create(jsp_codec codec, int width, int height, int bpp,
                             const uint8_t *palette, int palette_bytes, int device);
JSP_API void      jsp_destroy(jsp_dec *d);
JSP_API void      jsp_preinit(jsp_dec *d, int insignificant_lines);            /* IVideoCodec.Preinit        */
JSP_API int32_t  *jsp_previous_frame(jsp_dec *d);                               /* IVideoCodec.PreviousFrame  */
JSP_API int       jsp_is_key_frame(jsp_dec *d, const uint8_t *data, int len);   /* IVideoCodec.IsKeyFrame (host-side parse, no GPU) */
JSP_API jsp_state jsp_state_of(jsp_dec *d);                                     /* IVideoCodec.State          */
JSP_API jsp_state jsp_decompress_i(jsp_dec *d, const uint8_t *src, int len, int32_t *dst);   /* IVideoCodec.DecompressI */
JSP_API jsp_state jsp_continue_i(jsp_dec *d);                                   /* IVideoCodec.ContinueI (never in_progress, ScreenPressor.hx:210-215) */
JSP_API jsp_pframe_result jsp_decompress_p(jsp_dec *d, const uint8_t *src, int len, int32_t *dst); /* IVideoCodec.DecompressP */
JSP_API int       jsp_needs_index(jsp_dec *d);                                  /* IVideoCodec.NeedsIndex     */
JSP_API void      jsp_stop_and_clean(jsp_dec *d);                               /* IVideoCodec.StopAndClean   */

/* ---- batch path: many independent streams / GOPs per call (what the benchmarks drive) ----
 * One descriptor per stream; the frame table is what DataLoader keeps per stream
 * (DataLoader.hx:31 `frames`, VideoData.hx:68-73 CompressedFrame{key,data}). */
typedef struct {
    int32_t codec;               /* jsp_codec */
    int32_t width, height, bpp;
    const uint8_t *palette;      /* MSVC8 only */
    int32_t palette_bytes;
    int32_t n_frames;
    const uint8_t  *bytes;       /* compressed bytes of this stream (host; pinned for async copies) */
    const uint64_t *frame_off;   /* n_frames offsets into bytes */
    const uint32_t *frame_len;   /* n_frames lengths */
    const uint8_t  *frame_key;   /* n_frames: 1 = key frame (DecompressI), 0 = DecompressP (Manager.hx:505-512) */
    int32_t sp_version;          /* ScreenPressor segments cut out of a longer stream: the stream's coder version (2, 3, 4)
                                    as jsp_segment_stream() reports it, 0 = none yet.  The reference creates its entropy
                                    coder at the first coded key frame and keeps it (ScreenPressor.hx:160-162); a segment
                                    that starts with a flat key frame needs to know it.  0 for whole streams. */
    int32_t reserved;            /* 0 */
} jsp_stream_desc;

typedef struct jsp_batch jsp_batch;

enum {
    JSP_BATCH_SIGNIFICANCE = 1,  /* also compute PFrameResult.significant_changes exactly (extra previous-frame reads) */
    JSP_BATCH_NUMA_BIND    = 2,  /* jsp_batch_create moves the CALLING thread to the CPUs of the device's NUMA node and prefers
                                    that node for its later allocations (jsp_host_alloc from this thread): the staging memory of
                                    the end-to-end path then sits next to the GPU's PCIe root.  No effect on a
                                    single-node host or with JSP_NUMA_BIND=0 in the environment. */
    JSP_BATCH_DISPLAY      = 4,  /* fused display epilogue (Manager.fill_bitmap_data Manager.hx:363-381): the
                                    MSVideo1 decode kernel STORES its pictures as the Int32 view of canvas bytes R,G,B,A (alpha
                                    255) instead of 0x00RRGGBB -- 0 extra bytes per pixel; every download of the batch then
                                    delivers MSVideo1 pictures in that format (pixels nobody wrote: 0xFF000000).  ScreenPressor pictures stay
                                    0x00RRGGBB in HBM (its predictors read back what they wrote) and are converted by the
                                    separate pass of jsp_batch_download_display. */
    JSP_BATCH_DISPLAY_FLIP = 8,  /* with JSP_BATCH_DISPLAY: picture row y is stored in row height-1-y (Main.hx:318,946) */
};

/* Per-frame result flags written by jsp_batch_results(). */
enum {
    JSP_FRAME_CHANGED     = 1,   /* data_pnt == dst (the frame altered pixels)              */
    JSP_FRAME_SIGNIFICANT = 2,   /* PFrameResult.significant_changes                         */
    JSP_FRAME_ERROR       = 4,   /* DecoderState.error_occured / malformed bitstream         */
    JSP_FRAME_DIFFERS     = 8,   /* key frames: Manager.frames_differ_significantly (Manager.hx:392-421) -- what
                                    Manager.worker stores into CompressedFrame.significant_changes for I frames (:499-504) */
};

JSP_API jsp_batch *jsp_batch_create(int device, int insignificant_lines, int flags);
JSP_API void       jsp_batch_destroy(jsp_batch *b);
/* Builds the frame / GOP tables, (re)allocates device memory for bitstreams and output pictures.
 * Frame order of every output array: stream 0 frames, stream 1 frames, ...  Returns total frames or <0. */
JSP_API int64_t    jsp_batch_configure(jsp_batch *b, const jsp_stream_desc *streams, int n_streams);
JSP_API int        jsp_batch_upload(jsp_batch *b);      /* host bitstreams -> HBM (async on the batch's stream) */
JSP_API int        jsp_batch_run(jsp_batch *b);         /* decode kernels only; inputs and outputs stay in HBM   */
JSP_API int        jsp_batch_sync(jsp_batch *b);
/* out_frames[i] (host, i in output order) receives picture i: width*height int32 0x00RRGGBB, bitstream
 * row order; NULL entries are skipped.  flags[i] = JSP_FRAME_* bits. */
JSP_API int        jsp_batch_download(jsp_batch *b, int32_t *const *out_frames, uint8_t *flags);
JSP_API int        jsp_batch_results(jsp_batch *b, uint8_t *flags);             /* flags only */
/* The search behind Manager.SkipStills (Manager.hx:289-317, DataLoader.FindPossibleChange DataLoader.hx:239-252) over the
 * decoded batch: first frame >= from_frame of `stream` whose change is significant (key frames: JSP_FRAME_DIFFERS, P frames:
 * JSP_FRAME_SIGNIFICANT -- needs JSP_BATCH_SIGNIFICANCE), else the stream's last frame; -1 on error. */
JSP_API int64_t    jsp_batch_next_significant(jsp_batch *b, int stream, int64_t from_frame);
/* Display epilogue of the caller (Manager.fill_bitmap_data, Manager.hx:363-381): pictures are converted on the device
 * from 0x00RRGGBB to the Int32 view of canvas bytes R,G,B,A (alpha 255; ScreenPressor at 16 bpp: 0xFF000000 | c << 3),
 * optionally flipped vertically (the negative-Y matrix Main applies when drawing, Main.hx:318,946), then downloaded. */
enum { JSP_DISPLAY_FLIP = 1 };
JSP_API int        jsp_batch_download_display(jsp_batch *b, int32_t *const *out_frames, uint8_t *flags, int display_flags);
/* upload + run + download, chunked and double-buffered over PCIe (the end-to-end path). */
JSP_API int        jsp_batch_decode_host(jsp_batch *b, int32_t *const *out_frames, uint8_t *flags);
/* Opt-in end-to-end path for a host that keeps ONE picture per stream and updates it in place -- what a player holding
 * PreviousFrame() does (Manager.hx:470-477).  The batch is decoded on the device as usual; then, frame index by frame index,
 * only the 16x16 blocks of a picture that differ from the stream's previous picture cross PCIe (a stream's first frame sends
 * every block) and host threads patch them into stream_pictures[s] (width*height int32, one per stream, NULL = skip the
 * stream).  on_frame (may be NULL) is called once per frame, in frame order per stream, when stream_pictures[stream] holds
 * exactly that frame; flags as jsp_batch_results.  Exact: a block is sent iff one of its pixels differs.  The default
 * contract -- one whole picture per frame -- stays jsp_batch_decode_host. */
typedef void (*jsp_frame_fn)(void *user, int32_t stream, int32_t frame, const int32_t *picture, uint8_t flags);
JSP_API int        jsp_batch_decode_host_delta(jsp_batch *b, int32_t *const *stream_pictures, uint8_t *flags,
                                               jsp_frame_fn on_frame, void *user);
JSP_API uint64_t   jsp_batch_delta_bytes(jsp_batch *b);     /* bytes the last jsp_batch_decode_host_delta moved device -> host */
/* device pointer (as integer) of output picture i and of the output arena; for device-resident consumers */
JSP_API uint64_t   jsp_batch_device_frame(jsp_batch *b, int64_t i);
/* Times `iters` back-to-back jsp_batch_run() passes with CUDA events on the batch's own stream after
 * `warmup` untimed passes.  ms_total = whole region; kernel_ms[k] (may be NULL, k < JSP_N_KERNELS) =
 * summed device time of kernel class k over the timed passes, launches[k] = launch count. */
enum { JSP_K_MSV1_DECODE = 0, JSP_K_FRAME_COPY = 1, JSP_K_SP_ENTROPY_RC = 2, JSP_K_SP_ENTROPY_ANS = 3,
       JSP_K_SP_RECON = 4, JSP_K_SIGNIF = 5, JSP_K_SP_ENTROPY_MIXED = 6 /* range-coder and rANS jobs in one launch */,
       JSP_N_KERNELS = 8 };
JSP_API int        jsp_batch_time_runs(jsp_batch *b, int warmup, int iters, int flush_l2,
                                       float *ms_total, float *kernel_ms, int64_t *launches);
/* Algorithmic bytes one jsp_batch_run() moves: output store + compressed read +
 * previous-frame reads for copied pixels; in_bytes/out_bytes = PCIe bytes of the end-to-end path. */
JSP_API int        jsp_batch_stats(jsp_batch *b, uint64_t *pixels, uint64_t *alg_bytes,
                                   uint64_t *in_bytes, uint64_t *out_bytes);

/* bytes[k] (k < JSP_N_KERNELS) = algorithmic bytes kernel class k moves in one jsp_batch_run(): pictures it
 * writes (and previous pictures it copies) + compressed bytes it reads; intermediates excluded. */
JSP_API int        jsp_batch_kernel_bytes(jsp_batch *b, uint64_t *bytes);

/* Entropy-coded symbols the ScreenPressor kernels decoded in the last jsp_batch_run() (all frames); the entropy stage
 * is latency-bound, so its throughput is reported in symbols / s rather than bytes / s. */
JSP_API int64_t    jsp_batch_symbols(jsp_batch *b);

/* One-shot convenience with the signature SURVEY.md 8b sketches; shards streams longest-first over
 * n_gpus devices of this process (no collectives: GOPs/streams are independent). */
JSP_API int jsp_batch_decode(const jsp_stream_desc *streams, int n_streams, int n_gpus,
                             int32_t *const *out_frames, uint8_t *out_changed,
                             uint8_t *out_significant, int32_t *out_status);

/* ---- AVI indexer: a complete RIFF AVI file in memory -> video stream info + frame table ----
 * Local-file counterpart of AVIParser (AVIParser.hx:42-184) and of the idx1 / OpenDML index handling of the loaders
 * (DataLoaderAVIIndexed.hx:276-350, DataLoader.hx:321-401).  Frames are NOT copied: offsets point into `file`, so a
 * jsp_stream_desc built from the table uploads straight from the (ideally pinned) file buffer. */
typedef struct {
    int32_t codec;            /* jsp_codec chosen from the fourcc as AVIParser.hx:75-78 does */
    int32_t width, height, bpp;
    uint32_t fourcc;
    int32_t n_frames;         /* video chunks found in movi */
    int32_t n_frames_header;  /* avih dwTotalFrames */
    int32_t palette_bytes;    /* strf bytes from offset 40 (8 bpp) */
    int32_t has_index;        /* idx1 or OpenDML index present: key flags come from it */
    double  fps;
} jsp_avi_info;
typedef struct jsp_avi jsp_avi;
JSP_API jsp_avi    *jsp_avi_parse(const uint8_t *file, uint64_t size);          /* NULL on failure, see jsp_avi_last_error */
JSP_API void        jsp_avi_free(jsp_avi *a);
JSP_API int         jsp_avi_get_info(const jsp_avi *a, jsp_avi_info *out);
JSP_API int         jsp_avi_get_palette(const jsp_avi *a, uint8_t *out, int cap);  /* returns palette_bytes */
/* off[i] / len[i]: payload of video frame i inside `file`; key[i]: index key flag; key_known[i] = 0 when no index
 * entry covers the frame (the caller then asks jsp_is_key_frame, DataLoaderAVIIndexed.hx:182). Returns n_frames. */
JSP_API int         jsp_avi_frame_table(const jsp_avi *a, uint64_t *off, uint32_t *len, uint8_t *key, uint8_t *key_known);
JSP_API const char *jsp_avi_last_error(void);

/* Frame table -> keyframe-delimited segments (GOPs), the unit of multi-GPU sharding and of seeking (Manager.hx:244-249).
 * seg_first[k] = first frame of segment k (capacity n_frames), seg_sp_version[k] (may be NULL) = the value to put into
 * jsp_stream_desc.sp_version for that segment.  Returns the number of segments. */
JSP_API int jsp_segment_stream(int32_t codec, const uint8_t *bytes, const uint64_t *frame_off, const uint32_t *frame_len,
                               const uint8_t *frame_key, int32_t n_frames, int32_t *seg_first, int32_t *seg_sp_version);

#ifdef __cplusplus
}
#endif
#endif
