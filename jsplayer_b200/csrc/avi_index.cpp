// avi_index.cpp -- local-file AVI indexer: RIFF walk -> video stream info + frame table (offset, length, key).
//
// The host-side counterpart of the reference's AVIParser (src/AVIParser.hx:42-184) and of the index handling of
// DataLoaderAVIIndexed / DataLoader (src/DataLoaderAVIIndexed.hx:276-350 idx1, src/DataLoader.hx:321-401 OpenDML
// ix00), for files that are completely in memory: the reference's resumable parser combinators exist only because
// its data arrives incrementally over HTTP, so a single forward walk replaces them.
// The frame table it yields is exactly what jsp_stream_desc wants; frames stay where they are in the file buffer
// (pin the buffer with jsp_host_alloc and the batcher uploads straight from it).
//
// Format facts honoured (file:line in the reference):
//  * avih: microseconds per frame (0 -> 66666), total frames, width, height        AVIParser.hx:42-62
//  * strh 'vids': fccHandler, dwLength at +32                                        AVIParser.hx:155-156
//  * strf: bits per pixel at +14, biCompression at +16 when strh's handler is 0, palette from +40 for 8 bpp;
//    codec = MSVideo1 when the fourcc is MSVC / msvc / CRAM / 0, ScreenPressor otherwise   AVIParser.hx:64-88
//  * movi: 00dc / 00db chunks, also inside LIST 'rec '; chunks are padded to even sizes; the frame length handed to
//    the codec is the chunk's true size (the reference passes the pad byte along, AVIParser.hx:144)
//  * idx1: 16-byte records, video = id & 0xFF0000 == 0x640000, key = flags & 16, offsets relative to the 'movi'
//    fourcc unless the first offset lies beyond it (then absolute)                   DataLoaderAVIIndexed.hx:298-330
//  * OpenDML ix00 (in movi) / indx standard index (in strl): entries (offset, size), key = bit 31 of size clear,
//    offset 0 repeats the previous one                                               DataLoader.hx:321-361
//  * without any index the codec's own IsKeyFrame decides (DataLoaderAVIIndexed.hx:182) -- done by the caller.
#include "../../include/jsplayer_cuda.h"
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

struct Frame { uint64_t off; uint32_t len; uint8_t key; uint8_t key_known; };

inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
constexpr uint32_t fcc(const char (&s)[5]) { return (uint32_t)(uint8_t)s[0] | ((uint32_t)(uint8_t)s[1] << 8) | ((uint32_t)(uint8_t)s[2] << 16) | ((uint32_t)(uint8_t)s[3] << 24); }
inline bool is_video_id(uint32_t id) { return (id & 0xFF0000u) == 0x640000u; }      // '??d?' : 00dc / 00db

}  // namespace

struct jsp_avi {
    jsp_avi_info info{};
    std::vector<uint8_t> palette;
    std::vector<Frame> frames;            // in file order
    uint64_t movi_pos = 0;                // offset of the 'movi' fourcc
    bool have_index = false;
    bool in_vids = false;
    uint32_t strh_fourcc = 0;
    struct IdxEntry { uint64_t data_off; uint32_t size; bool key; };
    std::vector<IdxEntry> pending;      // index entries, applied once every frame chunk is known (an index may precede its frames)
};

namespace {

void add_frame(jsp_avi *a, uint64_t data_off, uint32_t len)
{
    a->frames.push_back(Frame{data_off, len, 0, 0});
}

// marks frame at file offset `data_off` (chunk payload) with the index's key flag.  Frames are collected by one forward
// walk of the file, so their offsets ascend: a binary search per entry (an index with N bogus entries over M frames
// costs N log M, not N * M -- untrusted input).
void apply_index_entry(jsp_avi *a, size_t &cursor, uint64_t data_off, uint32_t size, bool key)
{
    if (cursor == SIZE_MAX) { a->pending.push_back({data_off, size, key}); return; }
    size_t lo = 0, hi = a->frames.size();
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        if (a->frames[mid].off < data_off) lo = mid + 1; else hi = mid;
    }
    if (lo < a->frames.size() && a->frames[lo].off == data_off) {
        a->frames[lo].key = key ? 1 : 0; a->frames[lo].key_known = 1;
        if (size < a->frames[lo].len) a->frames[lo].len = size;
    }
}

void parse_strf(jsp_avi *a, const uint8_t *p, uint32_t n)
{
    if (n < 20) return;
    const int32_t w = (int32_t)rd32(p + 4), h = (int32_t)rd32(p + 8);
    if (w > 0) a->info.width = w;
    if (h != 0) a->info.height = h < 0 ? -h : h;
    a->info.bpp = rd16(p + 14);
    uint32_t four = a->strh_fourcc;
    if (four == 0) four = rd32(p + 16);                                   // AVIParser.hx:70-72
    a->info.fourcc = four;
    if (four == fcc("MSVC") || four == fcc("msvc") || four == fcc("CRAM") || four == 0)
        a->info.codec = a->info.bpp == 8 ? JSP_CODEC_MSVC8 : JSP_CODEC_MSVC16;     // :75-78
    else
        a->info.codec = JSP_CODEC_SCREENPRESSOR;
    if (a->info.bpp == 8 && n > 40) a->palette.assign(p + 40, p + n);     // :79-85
    a->info.palette_bytes = (int32_t)a->palette.size();
}

// OpenDML standard index chunk payload ('ix00' in movi, or 'indx' with 2 longs per entry): DataLoader.hx:321-361
void parse_std_index(jsp_avi *a, const uint8_t *p, uint32_t n, size_t &cursor)
{
    if (n < 24) return;
    const uint32_t longs = rd16(p), nent = rd32(p + 4), ckid = rd32(p + 8);
    if (longs != 2 || !is_video_id(ckid)) return;
    const uint64_t base = (uint64_t)rd32(p + 12) | ((uint64_t)rd32(p + 16) << 32);
    uint32_t last_off = 0;
    for (uint32_t i = 0; i < nent && 24 + (uint64_t)i * 8 + 8 <= n; i++) {
        uint32_t off = rd32(p + 24 + i * 8);
        const uint32_t size = rd32(p + 24 + i * 8 + 4);
        if (off == 0) off = last_off; else last_off = off;                // :341-342
        apply_index_entry(a, cursor, base + off, size & 0x7FFFFFFFu, (size & 0x80000000u) == 0);
        a->have_index = true;
    }
}

// walks the chunks of [pos, end); `depth` guards against malformed nesting
void walk(jsp_avi *a, const uint8_t *f, uint64_t size, uint64_t pos, uint64_t end, int depth, bool in_movi, size_t &cursor)
{
    while (pos + 8 <= end && pos + 8 <= size) {
        const uint32_t id = rd32(f + pos);
        const uint64_t csz = rd32(f + pos + 4);
        const uint64_t body = pos + 8;
        uint64_t avail = csz;
        if (body + avail > size) avail = size - body;                     // truncated file: keep what is there
        if (id == fcc("RIFF") || id == fcc("LIST")) {
            if (avail >= 4 && depth < 8) {
                const uint32_t kind = rd32(f + body);
                const bool movi = kind == fcc("movi");
                if (movi && a->movi_pos == 0) a->movi_pos = body;
                if (kind == fcc("strl")) { a->in_vids = false; }
                walk(a, f, size, body + 4, body + avail, depth + 1, in_movi || movi, cursor);
            }
        } else if (id == fcc("avih") && avail >= 40) {
            uint32_t us = rd32(f + body);
            if (us == 0) us = 66666;                                       // AVIParser.hx:58
            a->info.fps = 1000000.0 / us;
            a->info.n_frames_header = (int32_t)rd32(f + body + 16);
            a->info.width = (int32_t)rd32(f + body + 32);
            a->info.height = (int32_t)rd32(f + body + 36);
        } else if (id == fcc("strh") && avail >= 8) {
            a->in_vids = rd32(f + body) == fcc("vids") && a->info.bpp == 0;   // the first video stream only
            if (a->in_vids) a->strh_fourcc = rd32(f + body + 4);
        } else if (id == fcc("strf")) {
            if (a->in_vids) { parse_strf(a, f + body, (uint32_t)avail); a->in_vids = false; }
        } else if (in_movi && (id == fcc("00dc") || id == fcc("00db"))) {
            add_frame(a, body, (uint32_t)avail);
        } else if (id == fcc("ix00") || id == fcc("indx")) {
            parse_std_index(a, f + body, (uint32_t)avail, cursor);
        } else if (id == fcc("idx1")) {
            // DataLoaderAVIIndexed.hx:298-330
            const uint32_t nrec = (uint32_t)(avail >> 4);
            int64_t first_offset = -1;
            for (uint32_t i = 0; i < nrec; i++) {
                const uint8_t *r = f + body + (uint64_t)i * 16;
                if (first_offset < 0) first_offset = rd32(r + 8);
            }
            const uint64_t base = (first_offset >= 0 && (uint64_t)first_offset < a->movi_pos) ? a->movi_pos : 0;
            for (uint32_t i = 0; i < nrec; i++) {
                const uint8_t *r = f + body + (uint64_t)i * 16;
                if (!is_video_id(rd32(r))) continue;
                // the record points at the chunk header; the payload starts 8 bytes later
                apply_index_entry(a, cursor, base + rd32(r + 8) + 8, rd32(r + 12), (rd32(r + 4) & 16u) != 0);
                a->have_index = true;
            }
        }
        pos = body + ((csz + 1) & ~(uint64_t)1);                          // chunks are padded to even sizes
    }
}

thread_local char g_avi_err[256] = "";

}  // namespace

extern "C" {

jsp_avi *jsp_avi_parse(const uint8_t *file, uint64_t size)
{
    if (!file || size < 12 || rd32(file) != fcc("RIFF") || rd32(file + 8) != fcc("AVI ")) {
        snprintf(g_avi_err, sizeof g_avi_err, "not a RIFF AVI file");
        return nullptr;
    }
    jsp_avi *a = nullptr;
    try {                                 // no C++ exception may cross the C ABI
        a = new jsp_avi();
        a->info.fps = 15.0;
        size_t cursor = SIZE_MAX;         // collect index entries while walking ...
        // top level: RIFF 'AVI ' followed by optional RIFF 'AVIX' extension segments (OpenDML)
        walk(a, file, size, 0, size, 0, false, cursor);
        cursor = 0;                       // ... and apply them now that every frame chunk is known
        for (const jsp_avi::IdxEntry &e : a->pending) apply_index_entry(a, cursor, e.data_off, e.size, e.key);
        a->pending.clear();
    } catch (...) {
        delete a;
        snprintf(g_avi_err, sizeof g_avi_err, "out of memory while indexing");
        return nullptr;
    }
    a->info.n_frames = (int32_t)a->frames.size();
    a->info.has_index = a->have_index ? 1 : 0;
    if (a->info.width <= 0 || a->info.height <= 0 || a->info.bpp == 0) {
        snprintf(g_avi_err, sizeof g_avi_err, "no video stream format found");
        delete a;
        return nullptr;
    }
    return a;
}

void jsp_avi_free(jsp_avi *a) { delete a; }

int jsp_avi_get_info(const jsp_avi *a, jsp_avi_info *out)
{
    if (!a || !out) return -1;
    *out = a->info;
    return 0;
}

int jsp_avi_get_palette(const jsp_avi *a, uint8_t *out, int cap)
{
    if (!a) return -1;
    const int n = (int)a->palette.size() < cap ? (int)a->palette.size() : cap;
    if (out && n > 0) memcpy(out, a->palette.data(), (size_t)n);
    return (int)a->palette.size();
}

int jsp_avi_frame_table(const jsp_avi *a, uint64_t *off, uint32_t *len, uint8_t *key, uint8_t *key_known)
{
    if (!a) return -1;
    for (size_t i = 0; i < a->frames.size(); i++) {
        if (off) off[i] = a->frames[i].off;
        if (len) len[i] = a->frames[i].len;
        if (key) key[i] = a->frames[i].key;
        if (key_known) key_known[i] = a->frames[i].key_known;
    }
    return (int)a->frames.size();
}

const char *jsp_avi_last_error(void) { return g_avi_err; }

// Cuts one stream's frame table into independently decodable segments: a segment starts at frame 0 and at every key
// frame (the unit the reference restarts from when seeking, Manager.hx:244-249).  ScreenPressor keeps ONE piece of
// state across key frames: the entropy coder is created by the first coded key frame that names a known version and is
// never replaced (`if (ec == null) initEntro(version)`, ScreenPressor.hx:160-162), and a flat key frame needs it to
// exist (:112-114).  seg_sp_version[k] carries that state into segment k (-> jsp_stream_desc.sp_version), so that a
// segment which begins with a flat key frame, or whose key frame carries another version nibble, decodes exactly as
// it does in stream order.
int jsp_segment_stream(int32_t codec, const uint8_t *bytes, const uint64_t *frame_off, const uint32_t *frame_len,
                       const uint8_t *frame_key, int32_t n_frames, int32_t *seg_first, int32_t *seg_sp_version)
{
    if (n_frames < 0 || (n_frames > 0 && (!frame_key || !seg_first))) return -1;
    int n_seg = 0, version = 0;
    for (int32_t i = 0; i < n_frames; i++) {
        if (i == 0 || frame_key[i]) {
            seg_first[n_seg] = i;
            if (seg_sp_version) seg_sp_version[n_seg] = version;
            n_seg++;
        }
        if (codec == JSP_CODEC_SCREENPRESSOR && version == 0 && frame_key[i] && bytes && frame_off && frame_len && frame_len[i] > 0) {
            const int head = bytes[frame_off[i]], v = (head >> 4) + 1;
            if ((head & 0xF) == 2 && v >= 2 && v <= 4) version = v;       // initEntro succeeded: sticky from here on
        }
    }
    return n_seg;
}

}  // extern "C"
