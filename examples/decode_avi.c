/* decode_avi.c -- plain-C host of libjsplayer_cuda: AVI file -> keyframe-delimited segments -> batch decode on the GPU.
 *
 *   gcc -std=c99 -I include examples/decode_avi.c -L jsplayer_b200 -ljsplayer_cuda -Wl,-rpath,$PWD/jsplayer_b200 -o decode_avi
 *   ./decode_avi clip.avi [n_gpus]
 *
 * The same calls are what a Haxe/hxcpp (haxe/JsplayerCuda.hx), Go (cgo) or Java (JNI) host binds: plain pointers and
 * sizes only.  Prints one line per frame: index, key flag, changed / significant / differs flags, a checksum. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jsplayer_cuda.h"

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s file.avi [n_gpus]\n", argv[0]); return 2; }
    if (jsp_device_count() <= 0) { fprintf(stderr, "no CUDA device: libjsplayer_cuda has no CPU fallback\n"); return 3; }
    FILE *fh = fopen(argv[1], "rb");
    if (!fh) { perror(argv[1]); return 1; }
    fseek(fh, 0, SEEK_END);
    long size = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    uint8_t *file = (uint8_t *)jsp_host_alloc((size_t)size + 64);      /* pinned: the batcher uploads straight from it */
    if (!file || fread(file, 1, (size_t)size, fh) != (size_t)size) { fprintf(stderr, "read failed\n"); return 1; }
    fclose(fh);

    jsp_avi *avi = jsp_avi_parse(file, (uint64_t)size);
    if (!avi) { fprintf(stderr, "%s: %s\n", argv[1], jsp_avi_last_error()); return 1; }
    jsp_avi_info info;
    jsp_avi_get_info(avi, &info);
    int n = info.n_frames;
    uint64_t *off = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
    uint32_t *len = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
    uint8_t *key = (uint8_t *)malloc((size_t)n), *known = (uint8_t *)malloc((size_t)n);
    uint8_t palette[1024];
    int pal_bytes = jsp_avi_get_palette(avi, palette, (int)sizeof palette);
    jsp_avi_frame_table(avi, off, len, key, known);
    jsp_avi_free(avi);

    /* frames no index entry covers: ask the codec (DataLoaderAVIIndexed.hx:182); host-side parse, no GPU work */
    jsp_dec *probe = jsp_create((jsp_codec)info.codec, info.width, info.height, info.bpp, NULL, 0, -1);
    for (int i = 0; i < n; i++)
        if (!known[i]) key[i] = (uint8_t)jsp_is_key_frame(probe, file + off[i], (int)len[i]);
    jsp_destroy(probe);

    /* one descriptor per keyframe-delimited segment: the unit of sharding over GPUs (no exchange step exists) */
    int32_t *seg_first = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1)), *seg_ver = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int n_seg = jsp_segment_stream(info.codec, file, off, len, key, n, seg_first, seg_ver);
    if (n_seg < 0) { fprintf(stderr, "segmenting failed\n"); return 1; }
    jsp_stream_desc *sd = (jsp_stream_desc *)calloc((size_t)n_seg + 1, sizeof *sd);
    for (int k = 0; k < n_seg; k++) {
        const int i = seg_first[k];
        jsp_stream_desc *d = &sd[k];
        d->codec = info.codec; d->width = info.width; d->height = info.height; d->bpp = info.bpp;
        d->palette = pal_bytes > 0 ? palette : NULL; d->palette_bytes = pal_bytes > 1024 ? 1024 : pal_bytes;
        d->bytes = file; d->frame_off = off + i; d->frame_len = len + i; d->frame_key = key + i;
        d->n_frames = (k + 1 < n_seg ? seg_first[k + 1] : n) - i;
        d->sp_version = seg_ver[k];          /* ScreenPressor keeps its entropy coder across key frames */
    }
    size_t npix = (size_t)info.width * info.height;
    int32_t *pictures = (int32_t *)jsp_host_alloc(npix * 4 * (size_t)n);
    int32_t **out = (int32_t **)malloc(sizeof(int32_t *) * (size_t)n);
    for (int i = 0; i < n; i++) out[i] = pictures + npix * (size_t)i;
    uint8_t *changed = (uint8_t *)calloc((size_t)n, 1), *signif = (uint8_t *)calloc((size_t)n, 1);
    int32_t *status = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    int n_gpus = argc > 2 ? atoi(argv[2]) : 1;
    if (jsp_batch_decode(sd, n_seg, n_gpus, out, changed, signif, status) != 0) {
        fprintf(stderr, "decode failed: %s\n", jsp_last_error());
        return 1;
    }
    printf("%s: %dx%d %d bpp, %d frames in %d segments, codec %d\n", argv[1], info.width, info.height, info.bpp, n, n_seg, info.codec);
    for (int i = 0; i < n; i++) {
        uint32_t sum = 0;
        for (size_t p = 0; p < npix; p++) sum = sum * 31u + (uint32_t)out[i][p];
        printf("frame %4d key %d changed %d significant %d status %d checksum %08x\n", i, key[i], changed[i], signif[i], status[i], sum);
    }
    jsp_host_free(pictures); jsp_host_free(file);
    return 0;
}
