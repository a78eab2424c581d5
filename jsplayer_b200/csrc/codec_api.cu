// codec_api.cu -- the per-stream drop-in: one stateful decoder per stream with the members of
// `interface IVideoCodec` (reference src/IVideoCodec.hx:16-29), host frame buffers in, host frame
// buffers out, every picture decoded by the sm_100a kernels (a batch of one frame).
//
// Buffer protocol (reference src/Manager.hx:424-443,470-477): the caller owns all frame buffers; the
// codec remembers the last buffer whose picture changed as PreviousFrame() and the caller never passes
// that buffer as dst.  The device keeps its own copy of the previous picture, so the retained host
// buffer is only ever returned, never read.
#include "batch.cuh"
#include <cstring>
#include <vector>

using namespace jsp;

struct jsp_dec {
    int codec, w, h, bpp, device;
    std::vector<uint8_t> palette;
    jsp_batch *b = nullptr;
    int32_t *d_prev = nullptr;        // device copy of the previous picture
    bool d_prev_valid = false;
    int32_t *h_prev = nullptr;        // what PreviousFrame() returns (caller's buffer)
    uint8_t *h_src = nullptr;         // pinned staging for the compressed frame
    size_t h_src_cap = 0;
    int insign_lines = 0;
    bool preinit_done = false;
    jsp_state state = JSP_ZERO_STATE;
    // ScreenPressor bookkeeping (ScreenPressor.hx:41-43)
    bool decodedI = false;
};

namespace {

// MSVideo1.hx:226-259 / :395-427, host-side parse (no GPU), JavaScript `undefined` reads made explicit
int msv1_is_key(bool is8, int X, int Y, const uint8_t *src, int len)
{
    if (len == 0) return 0;
    const int nbx = X >> 2, nby = Y >> 2;
    long skip = 0; int si = 0, key = 1;
    for (int by = 0; by < nby; by++)
        for (int bx = 0; bx < nbx; bx++) {
            if (skip != 0) { skip--; continue; }
            const int a = si < len ? src[si] : -1, b = si + 1 < len ? src[si + 1] : -1;
            if (is8 && a >= 0 && b >= 0 && a + b == 0) return key;
            si += 2;
            if (b >= 0 && (b & 0xFC) == 0x84) {
                if (!is8) return 0;
                skip = ((b - 0x84) << 8) + a - 1; key = 0;
            } else if (b >= 0 && b < 0x80) {
                if (is8) si += 2;
                else si += (si + 1 < len && (src[si + 1] & 0x80)) ? 16 : 4;
            } else if (is8 && b >= 0x90)
                si += 8;
        }
    return key;
}

bool ensure_batch(jsp_dec *d)
{
    if (d->b) return true;
    d->b = jsp_batch_create(d->device, d->insign_lines, JSP_BATCH_SIGNIFICANCE);
    if (!d->b) return false;
    d->b->persist_streams = 1;
    const size_t npix = ((size_t)d->w * d->h + 63) & ~(size_t)63;
    if (!JSP_CUDA(cudaMalloc((void **)&d->d_prev, npix * 4))) return false;
    return JSP_CUDA(cudaMemset(d->d_prev, 0, npix * 4));
}

// Decodes one frame through the batch path. Returns false on error; *flags = JSP_FRAME_* bits.
bool decode_one(jsp_dec *d, const uint8_t *src, int len, int key, int32_t *dst, uint8_t *flags)
{
    if (!ensure_batch(d)) return false;
    jsp_batch *b = d->b;
    b->insign_lines = d->insign_lines;
    if ((size_t)len + 64 > d->h_src_cap) {
        if (d->h_src) cudaFreeHost(d->h_src);
        d->h_src = nullptr; d->h_src_cap = 0;
        const size_t cap = (size_t)len * 2 + 4096;
        if (!JSP_CUDA(cudaHostAlloc((void **)&d->h_src, cap, cudaHostAllocDefault))) return false;
        d->h_src_cap = cap;
    }
    if (len > 0) memcpy(d->h_src, src, (size_t)len);
    const uint64_t off = 0; const uint32_t l = (uint32_t)len; const uint8_t k = (uint8_t)key;
    jsp_stream_desc sd{};
    sd.codec = d->codec; sd.width = d->w; sd.height = d->h; sd.bpp = d->bpp;
    sd.palette = d->palette.empty() ? nullptr : d->palette.data(); sd.palette_bytes = (int)d->palette.size();
    sd.n_frames = 1; sd.bytes = d->h_src; sd.frame_off = &off; sd.frame_len = &l; sd.frame_key = &k;
    b->ext_prev = d->d_prev_valid ? d->d_prev : nullptr;
    b->ext_has_prev = d->h_prev != nullptr;
    if (jsp_batch_configure(b, &sd, 1) < 0) return false;
    int32_t *outs[1] = {dst};
    if (jsp_batch_decode_host(b, outs, flags)) return false;
    if (*flags & JSP_FRAME_CHANGED) {
        const size_t bytes = (size_t)d->w * d->h * 4;
        if (!JSP_CUDA(cudaMemcpyAsync(d->d_prev, reinterpret_cast<const void *>((uintptr_t)jsp_batch_device_frame(b, 0)), bytes,
                                      cudaMemcpyDeviceToDevice, b->st_compute))) return false;
        if (!JSP_CUDA(cudaStreamSynchronize(b->st_compute))) return false;
        d->d_prev_valid = true;
    }
    return true;
}

}  // namespace

extern "C" {

jsp_dec *jsp_create(jsp_codec codec, int width, int height, int bpp, const uint8_t *palette, int palette_bytes, int device)
{
    if (width <= 0 || height <= 0) { set_error("jsp_create: bad size %dx%d", width, height); return nullptr; }
    if (codec != JSP_CODEC_MSVC16 && codec != JSP_CODEC_MSVC8 && codec != JSP_CODEC_SCREENPRESSOR) {
        set_error("jsp_create: unknown codec %d", (int)codec); return nullptr;
    }
    jsp_dec *d = new jsp_dec();
    d->codec = codec; d->w = width; d->h = height; d->bpp = bpp; d->device = device;
    if (palette && palette_bytes > 0) d->palette.assign(palette, palette + palette_bytes);
    return d;     // the device is touched lazily: IsKeyFrame / NeedsIndex work without a GPU
}

void jsp_destroy(jsp_dec *d)
{
    if (!d) return;
    if (d->b) { cudaSetDevice(d->b->device); if (d->d_prev) cudaFree(d->d_prev); jsp_batch_destroy(d->b); }
    if (d->h_src) cudaFreeHost(d->h_src);
    delete d;
}

void jsp_preinit(jsp_dec *d, int insignificant_lines)
{
    if (!d) return;
    d->insign_lines = insignificant_lines; d->preinit_done = true;
}

int32_t *jsp_previous_frame(jsp_dec *d) { return d ? d->h_prev : nullptr; }

int jsp_is_key_frame(jsp_dec *d, const uint8_t *data, int len)
{
    if (!d) return 0;
    if (d->codec == JSP_CODEC_SCREENPRESSOR) {     // ScreenPressor.hx:96-101
        if (!data || len <= 0) return 0;
        const uint8_t b = data[0];
        return b == 0x12 || b == 0x11 || b == 0x22 || b == 0x21 || b == 0x32 || b == 0x31;
    }
    if (!data || len <= 0) return 0;
    return msv1_is_key(d->codec == JSP_CODEC_MSVC8, d->w, d->h, data, len);
}

jsp_state jsp_state_of(jsp_dec *d) { return d ? d->state : JSP_ERROR_OCCURED; }
int jsp_needs_index(jsp_dec *d) { return d && d->codec != JSP_CODEC_SCREENPRESSOR; }   // MSVideo1.hx:221-224, ScreenPressor.hx:486-489
jsp_state jsp_continue_i(jsp_dec *d) { return d ? d->state : JSP_ERROR_OCCURED; }      // never in_progress (ScreenPressor.hx:210-215)

jsp_pframe_result jsp_decompress_p(jsp_dec *d, const uint8_t *src, int len, int32_t *dst)
{
    jsp_pframe_result r{nullptr, 0};
    if (!d || !dst || len < 0 || (len > 0 && !src)) { set_error("jsp_decompress_p: bad arguments"); return r; }
    r.data_pnt = d->h_prev;
    if (d->codec == JSP_CODEC_SCREENPRESSOR) {
        // ScreenPressor.hx:308-313: nothing to do before the first I frame, for empty frames, or when
        // the change flag is 0 -- dst is not touched and the previous buffer is returned
        if (len == 0 || !d->decodedI || src[0] == 0) return r;
    }
    // MSVideo1.hx:109-110: empty / skip-only RGB555 frames return the previous buffer, dst untouched
    if (d->codec == JSP_CODEC_MSVC16 && msv16_unchanged(d->w, d->h, src, (uint32_t)len)) return r;
    uint8_t fl = 0;
    if (!decode_one(d, src, len, 0, dst, &fl)) { d->state = JSP_ERROR_OCCURED; return r; }
    if (fl & JSP_FRAME_ERROR) d->state = JSP_ERROR_OCCURED;
    if (fl & JSP_FRAME_CHANGED) d->h_prev = dst;
    r.data_pnt = d->h_prev;
    r.significant_changes = (fl & JSP_FRAME_SIGNIFICANT) ? 1 : 0;
    return r;
}

jsp_state jsp_decompress_i(jsp_dec *d, const uint8_t *src, int len, int32_t *dst)
{
    if (!d || !dst || len < 0 || (len > 0 && !src)) { set_error("jsp_decompress_i: bad arguments"); return JSP_ERROR_OCCURED; }
    if (d->codec != JSP_CODEC_SCREENPRESSOR) {     // MSVideo1.hx:62-67: DecompressI is DecompressP, result dropped
        jsp_decompress_p(d, src, len, dst);
        return d->state == JSP_ERROR_OCCURED ? JSP_ERROR_OCCURED : JSP_ZERO_STATE;
    }
    uint8_t fl = 0;
    if (!decode_one(d, src, len, 1, dst, &fl)) { d->state = JSP_ERROR_OCCURED; return JSP_ERROR_OCCURED; }
    if (fl & JSP_FRAME_ERROR) return JSP_ERROR_OCCURED;           // ScreenPressor.hx:157-162
    d->h_prev = dst; d->decodedI = true;                          // ScreenPressor.hx:290-293
    return JSP_ZERO_STATE;
}

void jsp_stop_and_clean(jsp_dec *d)
{
    if (!d) return;                                               // ScreenPressor.hx:81-84
    d->h_prev = nullptr; d->d_prev_valid = false; d->decodedI = false;
}

}  // extern "C"
