/*
 * vf_raisr_replay.c -- a plain-C caller of libraisr.so that replays what ffmpeg/vf_raisr.c does to the library, without FFmpeg:
 *   init()              vf_raisr.c:98-156   range/asm strings -> RNLHandler_Init
 *   config_props_*      vf_raisr.c:180-224  AV_CEIL_RSHIFT plane geometry, out = in * ratio, evenoutput
 *   filter_frame()      vf_raisr.c:226-332  fresh output planes per frame, RNLHandler_SetRes on frame 0, RNLHandler_Process per frame
 *   uninit()            vf_raisr.c:334-337  RNLHandler_Deinit
 * Frame memory is what av_frame_get_buffer hands out: pageable malloc'ed planes whose linesize is the row size rounded up to
 * 64 bytes plus padding (libavutil/frame.c get_video_buffer), so rows are NOT contiguous.
 *
 * Compiled as C (not C++) against include/raisr/RaisrHandler.h: also the proof that the public headers are C-clean.
 *
 * usage: vf_raisr_replay <filterfolder> <w> <h> <bits> <ratio> <passes> <mode> <blending> <pixfmt: 420|422|444> <evenoutput>
 *                        <frames> <in.yuv> <out.yuv>
 * in.yuv holds <frames> planar frames (Y, U, V; 16-bit little endian samples when bits > 8), out.yuv receives the outputs.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "raisr/RaisrDefaults.h"
#include "raisr/RaisrHandler.h"

#define CEIL_RSHIFT(a, b) (-((-(a)) >> (b)))             /* AV_CEIL_RSHIFT */
#define ALIGN_UP(x, a) (((x) + (a) - 1) & ~((a) - 1))     /* FFALIGN */

struct plane {
    unsigned char *base, *data;
    int width, height, linesize;
};

/* one plane the way get_video_buffer lays it out: linesize aligned to 64, a few padding rows, data offset into the allocation */
static int plane_alloc(struct plane *p, int w, int h, int bps)
{
    p->width = w;
    p->height = h;
    p->linesize = ALIGN_UP(w * bps, 64) + 64;
    p->base = (unsigned char *)malloc((size_t)p->linesize * (h + 32) + 64);
    if (!p->base) return -1;
    memset(p->base, 0xA5, (size_t)p->linesize * (h + 32) + 64);      /* padding must never leak into the result */
    p->data = p->base + 64 - ((uintptr_t)p->base & 63);
    return 0;
}

static void plane_free(struct plane *p) { free(p->base); p->base = NULL; }

static int plane_read(struct plane *p, int bps, FILE *f)
{
    for (int y = 0; y < p->height; y++)
        if (fread(p->data + (size_t)y * p->linesize, bps, p->width, f) != (size_t)p->width) return -1;
    return 0;
}

static int plane_write(const struct plane *p, int bps, FILE *f)
{
    for (int y = 0; y < p->height; y++)
        if (fwrite(p->data + (size_t)y * p->linesize, bps, p->width, f) != (size_t)p->width) return -1;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc != 14) {
        fprintf(stderr, "usage: %s filterfolder w h bits ratio passes mode blending pixfmt evenoutput frames in.yuv out.yuv\n", argv[0]);
        return 2;
    }
    const char *folder = argv[1];
    const int w = atoi(argv[2]), h = atoi(argv[3]), bits = atoi(argv[4]);
    const float ratio = (float)atof(argv[5]);
    const int passes = atoi(argv[6]), mode = atoi(argv[7]), blending = atoi(argv[8]), pixfmt = atoi(argv[9]);
    const int evenoutput = atoi(argv[10]), frames = atoi(argv[11]);
    const int bps = bits > 8 ? 2 : 1;
    const int hsub = pixfmt == 444 ? 0 : 1, vsub = pixfmt == 420 ? 1 : 0;        /* log2_chroma_w / log2_chroma_h */

    /* init(): vf_raisr.c:146 -- threadcount and asm are CPU notions, passed through like the filter does */
    RNLERRORTYPE ret = RNLHandler_Init(folder, ratio, (unsigned)bits, VideoRange, 20, AVX512_FP16, (unsigned)passes, (unsigned)mode);
    if (ret != RNLErrorNone) {
        fprintf(stderr, "RNLHandler_Init error 0x%08x\n", (unsigned)ret);
        return 3;
    }

    /* config_props_output(): vf_raisr.c:207-224 */
    int out_w = (int)(w * ratio), out_h = (int)(h * ratio);
    if (evenoutput == 1) {
        out_w -= out_w % 2;
        out_h -= out_h % 2;
    }

    FILE *fin = fopen(argv[12], "rb"), *fout = fopen(argv[13], "wb");
    if (!fin || !fout) {
        fprintf(stderr, "cannot open the yuv files\n");
        return 4;
    }

    struct plane in[3];
    for (int p = 0; p < 3; p++)                                                   /* config_props_input(): vf_raisr.c:180-205 */
        if (plane_alloc(&in[p], p ? CEIL_RSHIFT(w, hsub) : w, p ? CEIL_RSHIFT(h, vsub) : h, bps)) return 5;

    int rc = 0;
    for (int n = 0; n < frames && rc == 0; n++) {
        struct plane out[3];
        VideoDataType vdt_in[3], vdt_out[3];
        memset(vdt_in, 0, sizeof(vdt_in));
        memset(vdt_out, 0, sizeof(vdt_out));
        for (int p = 0; p < 3; p++) {
            if (plane_read(&in[p], bps, fin)) { fprintf(stderr, "short input\n"); return 6; }
            /* ff_get_video_buffer(): a fresh buffer per frame (the pool recycles a handful of them) */
            if (plane_alloc(&out[p], p ? CEIL_RSHIFT(out_w, hsub) : out_w, p ? CEIL_RSHIFT(out_h, vsub) : out_h, bps)) return 5;
            vdt_in[p].pData = in[p].data;   vdt_in[p].width = in[p].width;   vdt_in[p].height = in[p].height;   vdt_in[p].step = in[p].linesize;
            vdt_out[p].pData = out[p].data; vdt_out[p].width = out[p].width; vdt_out[p].height = out[p].height; vdt_out[p].step = out[p].linesize;
        }
        if (n == 0) {                                                             /* vf_raisr.c:286-302 */
            ret = RNLHandler_SetRes(&vdt_in[0], &vdt_in[1], &vdt_in[2], &vdt_out[0], &vdt_out[1], &vdt_out[2]);
            if (ret != RNLErrorNone) { fprintf(stderr, "RNLHandler_SetRes error 0x%08x\n", (unsigned)ret); rc = 7; }
        }
        if (rc == 0) {                                                            /* vf_raisr.c:305-318 */
            ret = RNLHandler_Process(&vdt_in[0], &vdt_in[1], &vdt_in[2], &vdt_out[0], &vdt_out[1], &vdt_out[2], (BlendingMode)blending);
            if (ret != RNLErrorNone) { fprintf(stderr, "RNLHandler_Process error 0x%08x\n", (unsigned)ret); rc = 8; }
        }
        for (int p = 0; p < 3 && rc == 0; p++) {
            if (plane_write(&out[p], bps, fout)) rc = 9;
            /* the bytes right of each row and below the plane are the caller's: the library must not have written there */
            for (int y = 0; y < out[p].height && rc == 0; y++)
                for (int x = out[p].width * bps; x < out[p].linesize; x++)
                    if (out[p].data[(size_t)y * out[p].linesize + x] != 0xA5) { fprintf(stderr, "padding overwritten (plane %d row %d)\n", p, y); rc = 10; break; }
        }
        for (int p = 0; p < 3; p++) plane_free(&out[p]);
    }
    RNLHandler_Deinit();                                                          /* uninit(): vf_raisr.c:334-337 */
    for (int p = 0; p < 3; p++) plane_free(&in[p]);
    fclose(fin);
    fclose(fout);
    return rc;
}
