/*
 * raisr_cuda -- RAISR super resolution on CUDA hardware frames (AV_PIX_FMT_CUDA) for the B200 engine (libraisr.so).
 *
 * The CUDA counterpart of the reference's OpenCL hardware-frame filter (ffmpeg/vf_raisr_opencl.c: filter_frame :70-154,
 * config_output :178-214, options :225-242) with the option surface of the software filter (ffmpeg/vf_raisr.c:81-94) incl.
 * `evenoutput` (vf_raisr.c:217-221).  Frames never leave the GPU: NVDEC -> raisr_cuda -> NVENC.
 *
 *   ffmpeg -init_hw_device cuda=cu:0 -filter_hw_device cu -hwaccel cuda -hwaccel_output_format cuda -i in.mp4 \
 *          -vf "raisr_cuda=ratio=2:filterfolder=/path/filters_2x/filters_lowres" -c:v hevc_nvenc out.mp4
 * (NVDEC's nv12 / p010 frames are taken as they are; planar yuv420p/422p/444p frames too)
 *
 * Instead of the process-global RNLHandler_* engine this filter owns a raisr_cuda_engine (include/raisr_cuda.h), so several
 * filter instances (and devices) can live in one process.  The engine is created with device = RAISR_CUDA_DEVICE_CALLER_CONTEXT:
 * it runs in the CUcontext FFmpeg made current (pushed around every call below) and never switches device itself, so the frames'
 * device pointers are valid for its kernels whether or not the AVCUDADeviceContext uses the primary context.
 *
 * Build: copy to libavfilter/, apply ffmpeg/0003-libavfilter-raisr_cuda.patch (Makefile/allfilters/configure lines), configure
 * with --enable-cuda-nvcc or --enable-ffnvcodec plus --enable-libraisr-cuda, link -lraisr.  NOT compiled in this repository (no
 * libav headers in the image); tests/harness/vf_raisr_replay.c exercises the same library calls against libraisr.so.
 *
 * This file follows FFmpeg's filter conventions and is meant to be contributed under LGPL 2.1+ like its siblings.
 */

#include "raisr_cuda.h"
#include "raisr/RaisrDefaults.h"

#include "libavutil/common.h"
#include "libavutil/hwcontext.h"
#include "libavutil/hwcontext_cuda_internal.h"
#include "libavutil/cuda_check.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"

#include "avfilter.h"
#include "internal.h"
#include "video.h"

#define CHECK_CU(x) FF_CUDA_CHECK_DL(avctx, s->hwctx->internal->cuda_dl, x)

typedef struct RaisrCudaContext {
    const AVClass *class;

    /* options: same names, ranges and defaults as vf_raisr.c:81-94 / vf_raisr_opencl.c:225-242 */
    float ratio;
    int bits;
    int range;
    char *filterfolder;
    int blending;
    int passes;
    int mode;
    int evenoutput;
    int numerics;

    AVCUDADeviceContext *hwctx;
    AVBufferRef *frames_ctx;            /* output frames */
    enum AVPixelFormat sw_format;
    raisr_cuda_engine *engine;
    int layout;                         /* 1 planar, 2 semi-planar (NV12 / P010) */
    int res_set;
} RaisrCudaContext;

static av_cold int raisr_cuda_init(AVFilterContext *avctx)
{
    return 0;                           /* the engine needs the device: created in config_output */
}

static av_cold void raisr_cuda_uninit(AVFilterContext *avctx)
{
    RaisrCudaContext *s = avctx->priv;

    if (s->engine && s->hwctx) {
        CUcontext dummy;
        CudaFunctions *cu = s->hwctx->internal->cuda_dl;
        CHECK_CU(cu->cuCtxPushCurrent(s->hwctx->cuda_ctx));
        raisr_cuda_destroy(s->engine);
        CHECK_CU(cu->cuCtxPopCurrent(&dummy));
        s->engine = NULL;
    }
    av_buffer_unref(&s->frames_ctx);
}

/* 1: planar YUV (three planes: yuv420p/422p/444p, 8 bits in bytes or 10 bits in little-endian 16-bit words);
 * 2: semi-planar 4:2:0 (NV12, P010: U and V interleaved in plane 1, P010 with its value in the high bits) -- the formats of
 *    vf_raisr_opencl.c:166-169; 0: not supported */
static int format_layout(const AVPixFmtDescriptor *desc, int bits)
{
    if (!desc || (desc->flags & (AV_PIX_FMT_FLAG_RGB | AV_PIX_FMT_FLAG_PAL | AV_PIX_FMT_FLAG_BITSTREAM | AV_PIX_FMT_FLAG_BE)))
        return 0;
    if (!(desc->flags & AV_PIX_FMT_FLAG_PLANAR) || desc->nb_components != 3 || desc->comp[0].depth != bits)
        return 0;
    if (desc->comp[0].plane == 0 && desc->comp[1].plane == 1 && desc->comp[2].plane == 2 &&
        desc->comp[0].shift == 0 && desc->comp[1].shift == 0 && desc->comp[2].shift == 0)
        return 1;
    if (desc->comp[0].plane == 0 && desc->comp[1].plane == 1 && desc->comp[2].plane == 1 &&
        desc->log2_chroma_w == 1 && desc->log2_chroma_h == 1 && desc->comp[1].offset < desc->comp[2].offset &&
        desc->comp[1].shift == desc->comp[0].shift && desc->comp[2].shift == desc->comp[0].shift)
        return 2;                       /* NV12 / P010 (U first); NV21 is not */
    return 0;
}

static int raisr_cuda_config_output(AVFilterLink *outlink)
{
    AVFilterContext *avctx = outlink->src;
    AVFilterLink *inlink = avctx->inputs[0];
    RaisrCudaContext *s = avctx->priv;
    AVHWFramesContext *in_frames, *out_frames;
    const AVPixFmtDescriptor *desc;
    CudaFunctions *cu;
    CUcontext dummy;
    raisr_cuda_config cfg = { 0 };
    int err, rc;

    if (!inlink->hw_frames_ctx) {
        av_log(avctx, AV_LOG_ERROR, "raisr_cuda needs CUDA hardware frames on its input\n");
        return AVERROR(EINVAL);
    }
    in_frames = (AVHWFramesContext *)inlink->hw_frames_ctx->data;
    if (in_frames->format != AV_PIX_FMT_CUDA)
        return AVERROR(EINVAL);
    s->hwctx = in_frames->device_ctx->hwctx;
    s->sw_format = in_frames->sw_format;
    desc = av_pix_fmt_desc_get(s->sw_format);
    if (desc && desc->comp[0].depth != s->bits) {              /* vf_raisr_opencl.c:193-197 */
        av_log(avctx, AV_LOG_ERROR, "input pixel doesn't match model's bitdepth\n");
        return AVERROR(EINVAL);
    }
    s->layout = format_layout(desc, s->bits);
    if (!s->layout) {
        av_log(avctx, AV_LOG_ERROR, "unsupported sw format %s: nv12, p010, or planar yuv420p/422p/444p (8 bit / 10-bit LE)\n",
               av_get_pix_fmt_name(s->sw_format));
        return AVERROR(ENOSYS);
    }

    outlink->w = inlink->w * s->ratio;                          /* vf_raisr.c:213-221 */
    outlink->h = inlink->h * s->ratio;
    if (s->evenoutput == 1) {
        outlink->w -= outlink->w % 2;
        outlink->h -= outlink->h % 2;
    }

    /* output frame pool on the same device (the pattern of vf_scale_cuda.c: init_hwframe_ctx) */
    av_buffer_unref(&s->frames_ctx);
    s->frames_ctx = av_hwframe_ctx_alloc(in_frames->device_ref);
    if (!s->frames_ctx)
        return AVERROR(ENOMEM);
    out_frames = (AVHWFramesContext *)s->frames_ctx->data;
    out_frames->format    = AV_PIX_FMT_CUDA;
    out_frames->sw_format = s->sw_format;
    out_frames->width     = FFALIGN(outlink->w, 32);
    out_frames->height    = FFALIGN(outlink->h, 32);
    err = av_hwframe_ctx_init(s->frames_ctx);
    if (err < 0)
        return err;
    av_buffer_unref(&outlink->hw_frames_ctx);
    outlink->hw_frames_ctx = av_buffer_ref(s->frames_ctx);
    if (!outlink->hw_frames_ctx)
        return AVERROR(ENOMEM);

    /* the engine: model tables + launch plan, in FFmpeg's CUDA context (vf_raisr_opencl.c:49-68 does the same for OpenCL) */
    cu = s->hwctx->internal->cuda_dl;
    err = CHECK_CU(cu->cuCtxPushCurrent(s->hwctx->cuda_ctx));
    if (err < 0)
        return err;
    cfg.model_path    = s->filterfolder;
    cfg.ratio         = s->ratio;
    cfg.bit_depth     = s->bits;
    cfg.range_type    = s->range;
    cfg.passes        = s->passes;
    cfg.two_pass_mode = s->mode;
    cfg.device        = RAISR_CUDA_DEVICE_CALLER_CONTEXT;
    cfg.numerics      = s->numerics;
    if (s->engine) {
        raisr_cuda_destroy(s->engine);
        s->engine = NULL;
    }
    rc = raisr_cuda_create(&cfg, &s->engine);
    CHECK_CU(cu->cuCtxPopCurrent(&dummy));
    if (rc != RNLErrorNone) {
        av_log(avctx, AV_LOG_ERROR, "raisr_cuda_create failed (0x%08x)\n", (unsigned)rc);
        return AVERROR(ENAVAIL);
    }
    s->res_set = 0;
    return 0;
}

static int raisr_cuda_filter_frame(AVFilterLink *inlink, AVFrame *in)
{
    AVFilterContext *avctx = inlink->dst;
    AVFilterLink *outlink = avctx->outputs[0];
    RaisrCudaContext *s = avctx->priv;
    const AVPixFmtDescriptor *desc = av_pix_fmt_desc_get(s->sw_format);
    CudaFunctions *cu = s->hwctx->internal->cuda_dl;
    CUcontext dummy;
    AVFrame *out = NULL;
    int err, rc;

    if (!in->hw_frames_ctx) {                                   /* vf_raisr_opencl.c:86-87 */
        err = AVERROR(EINVAL);
        goto fail;
    }
    out = av_frame_alloc();
    if (!out) {
        err = AVERROR(ENOMEM);
        goto fail;
    }
    err = av_hwframe_get_buffer(s->frames_ctx, out, 0);         /* pitched device planes */
    if (err < 0)
        goto fail;
    out->width  = outlink->w;
    out->height = outlink->h;

    err = CHECK_CU(cu->cuCtxPushCurrent(s->hwctx->cuda_ctx));
    if (err < 0)
        goto fail;
    if (!s->res_set) {                                          /* first frame: plane geometry (vf_raisr.c:286-302) */
        const int cw_in  = AV_CEIL_RSHIFT(in->width,  desc->log2_chroma_w), ch_in  = AV_CEIL_RSHIFT(in->height,  desc->log2_chroma_h);
        const int cw_out = AV_CEIL_RSHIFT(out->width, desc->log2_chroma_w), ch_out = AV_CEIL_RSHIFT(out->height, desc->log2_chroma_h);
        rc = raisr_cuda_set_res(s->engine, in->width, in->height, out->width, out->height, cw_in, ch_in, cw_out, ch_out);
        if (rc != RNLErrorNone) {
            CHECK_CU(cu->cuCtxPopCurrent(&dummy));
            av_log(avctx, AV_LOG_ERROR, "raisr_cuda_set_res error (0x%08x)\n", (unsigned)rc);
            err = AVERROR(ENOMEM);
            goto fail;
        }
        s->res_set = 1;
    }
    /* one launch per pass carries the whole frame (luma pass + both chroma resizes), asynchronous on FFmpeg's stream:
     * downstream CUDA consumers (NVENC, hwdownload) are ordered behind it on the same stream */
    if (s->layout == 2)                                         /* NV12 / P010 straight from NVDEC, straight into NVENC */
        rc = raisr_cuda_process_device_semiplanar(s->engine, in->data[0], in->linesize[0], in->data[1], in->linesize[1],
                                                  out->data[0], out->linesize[0], out->data[1], out->linesize[1],
                                                  desc->comp[0].shift, s->blending, s->hwctx->stream);
    else
        rc = raisr_cuda_process_device(s->engine,
                                       in->data[0],  in->linesize[0],  in->data[1],  in->linesize[1],  in->data[2],  in->linesize[2],
                                       out->data[0], out->linesize[0], out->data[1], out->linesize[1], out->data[2], out->linesize[2],
                                       s->blending, s->hwctx->stream);
    CHECK_CU(cu->cuCtxPopCurrent(&dummy));
    if (rc != RNLErrorNone) {
        av_log(avctx, AV_LOG_ERROR, "raisr_cuda_process_device error (0x%08x)\n", (unsigned)rc);
        err = AVERROR_EXTERNAL;
        goto fail;
    }

    err = av_frame_copy_props(out, in);
    if (err < 0)
        goto fail;
    av_frame_free(&in);
    return ff_filter_frame(outlink, out);

fail:
    av_frame_free(&in);
    av_frame_free(&out);
    return err;
}

#define OFFSET(x) offsetof(RaisrCudaContext, x)
#define FLAGS (AV_OPT_FLAG_FILTERING_PARAM | AV_OPT_FLAG_VIDEO_PARAM)
static const AVOption raisr_cuda_options[] = {
    {"ratio", "ratio of the upscaling, between 1 and 2", OFFSET(ratio), AV_OPT_TYPE_FLOAT, {.dbl = 2}, 1, 2, FLAGS},
    {"bits", "bit depth", OFFSET(bits), AV_OPT_TYPE_INT, {.i64 = 8}, 8, 10, FLAGS},
    {"range", "input color range", OFFSET(range), AV_OPT_TYPE_INT, {.i64 = VideoRange}, VideoRange, FullRange, FLAGS, "range"},
        { "video", NULL, 0, AV_OPT_TYPE_CONST, { .i64 = VideoRange }, INT_MIN, INT_MAX, FLAGS, "range" },
        { "full",  NULL, 0, AV_OPT_TYPE_CONST, { .i64 = FullRange },  INT_MIN, INT_MAX, FLAGS, "range" },
    {"filterfolder", "absolute filter folder path", OFFSET(filterfolder), AV_OPT_TYPE_STRING, {.str = "filters_2x/filters_lowres"}, 0, 0, FLAGS},
    {"blending", "CT blending mode (1: Randomness, 2: CountOfBitsChanged)", OFFSET(blending), AV_OPT_TYPE_INT,
     {.i64 = CountOfBitsChanged}, Randomness, CountOfBitsChanged, FLAGS, "blending"},
        { "Randomness",         NULL, 0, AV_OPT_TYPE_CONST, { .i64 = Randomness },         INT_MIN, INT_MAX, FLAGS, "blending" },
        { "CountOfBitsChanged", NULL, 0, AV_OPT_TYPE_CONST, { .i64 = CountOfBitsChanged }, INT_MIN, INT_MAX, FLAGS, "blending" },
    {"passes", "passes to run (1: one pass, 2: two pass)", OFFSET(passes), AV_OPT_TYPE_INT, {.i64 = 1}, 1, 2, FLAGS},
    {"mode", "mode for two pass (1: upscale in 1st pass, 2: upscale in 2nd pass)", OFFSET(mode), AV_OPT_TYPE_INT, {.i64 = 1}, 1, 2, FLAGS},
    {"evenoutput", "make output size as even number (0: ignore, 1: subtract 1px if needed)", OFFSET(evenoutput), AV_OPT_TYPE_INT, {.i64 = 0}, 0, 1, FLAGS},
    {"numerics", "hash numerics (0: IEEE, 1: bit-identical to the x86 AVX-512 build, 2: 1 when available, 3: fp16 filter stage)",
     OFFSET(numerics), AV_OPT_TYPE_INT, {.i64 = RAISR_NUMERICS_X86_IF_AVAILABLE}, 0, 3, FLAGS},
    {NULL}
};

AVFILTER_DEFINE_CLASS(raisr_cuda);

static const AVFilterPad raisr_cuda_inputs[] = {
    {
        .name         = "default",
        .type         = AVMEDIA_TYPE_VIDEO,
        .filter_frame = raisr_cuda_filter_frame,
    }
};

static const AVFilterPad raisr_cuda_outputs[] = {
    {
        .name         = "default",
        .type         = AVMEDIA_TYPE_VIDEO,
        .config_props = raisr_cuda_config_output,
    }
};

const AVFilter ff_vf_raisr_cuda = {
    .name           = "raisr_cuda",
    .description    = NULL_IF_CONFIG_SMALL("RAISR super resolution on CUDA frames (B200 engine)"),
    .priv_size      = sizeof(RaisrCudaContext),
    .priv_class     = &raisr_cuda_class,
    .init           = raisr_cuda_init,
    .uninit         = raisr_cuda_uninit,
    FILTER_INPUTS(raisr_cuda_inputs),
    FILTER_OUTPUTS(raisr_cuda_outputs),
    FILTER_SINGLE_PIXFMT(AV_PIX_FMT_CUDA),
    .flags_internal = FF_FILTER_FLAG_HWFRAME_AWARE,
};
