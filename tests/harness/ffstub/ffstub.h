/*
 * ffstub.h -- the handful of FFmpeg declarations ffmpeg/vf_raisr_cuda.c uses, restated from the public FFmpeg 6.x/7.x headers
 * (names, member names and signatures only) so that the filter source can be type-checked in an image without libav.
 * TEST INFRASTRUCTURE: `gcc -fsyntax-only`; nothing here is linked or shipped.
 */
#ifndef RAISR_FFSTUB_H
#define RAISR_FFSTUB_H
#include <limits.h>
#include <stddef.h>
#include <stdint.h>

#define av_cold
#define AVERROR(e) (-(e))
#define AVERROR_EXTERNAL (-0x20545845)
#define FFALIGN(x, a) (((x) + (a) - 1) & ~((a) - 1))
#define AV_CEIL_RSHIFT(a, b) (-((-(a)) >> (b)))
#define NULL_IF_CONFIG_SMALL(x) x
#define AV_LOG_ERROR 16
#include <errno.h>
#ifndef ENAVAIL
#define ENAVAIL 119
#endif
void av_log(void *avcl, int level, const char *fmt, ...);

enum AVPixelFormat { AV_PIX_FMT_NONE = -1, AV_PIX_FMT_YUV420P, AV_PIX_FMT_CUDA = 117 };
enum AVMediaType { AVMEDIA_TYPE_VIDEO };
#define AV_PIX_FMT_FLAG_BE (1 << 0)
#define AV_PIX_FMT_FLAG_PAL (1 << 1)
#define AV_PIX_FMT_FLAG_BITSTREAM (1 << 2)
#define AV_PIX_FMT_FLAG_PLANAR (1 << 4)
#define AV_PIX_FMT_FLAG_RGB (1 << 5)
typedef struct AVComponentDescriptor { int plane, step, offset, shift, depth; } AVComponentDescriptor;
typedef struct AVPixFmtDescriptor {
    const char *name; uint8_t nb_components, log2_chroma_w, log2_chroma_h; uint64_t flags; AVComponentDescriptor comp[4];
} AVPixFmtDescriptor;
const AVPixFmtDescriptor *av_pix_fmt_desc_get(enum AVPixelFormat pix_fmt);
const char *av_get_pix_fmt_name(enum AVPixelFormat pix_fmt);

typedef struct AVBufferRef { void *buffer; uint8_t *data; size_t size; } AVBufferRef;
AVBufferRef *av_buffer_ref(const AVBufferRef *buf);
void av_buffer_unref(AVBufferRef **buf);

typedef struct AVFrame {
    uint8_t *data[8]; int linesize[8]; int width, height, format; int64_t pts; AVBufferRef *hw_frames_ctx;
} AVFrame;
AVFrame *av_frame_alloc(void);
void av_frame_free(AVFrame **frame);
int av_frame_copy_props(AVFrame *dst, const AVFrame *src);

typedef struct AVHWDeviceContext { const void *av_class; int type; void *hwctx; } AVHWDeviceContext;
typedef struct AVHWFramesContext {
    const void *av_class; AVBufferRef *device_ref; AVHWDeviceContext *device_ctx; void *hwctx;
    int initial_pool_size; enum AVPixelFormat format, sw_format; int width, height;
} AVHWFramesContext;
AVBufferRef *av_hwframe_ctx_alloc(AVBufferRef *device_ctx);
int av_hwframe_ctx_init(AVBufferRef *ref);
int av_hwframe_get_buffer(AVBufferRef *hwframe_ctx, AVFrame *frame, int flags);

/* hwcontext_cuda_internal.h / dynlink_loader.h / cuda_check.h */
typedef struct CUctx_st *CUcontext;
typedef struct CUstream_st *CUstream;
typedef int CUresult;
typedef struct CudaFunctions {
    CUresult (*cuCtxPushCurrent)(CUcontext);
    CUresult (*cuCtxPopCurrent)(CUcontext *);
} CudaFunctions;
typedef struct AVCUDADeviceContextInternal { CudaFunctions *cuda_dl; } AVCUDADeviceContextInternal;
typedef struct AVCUDADeviceContext { CUcontext cuda_ctx; CUstream stream; AVCUDADeviceContextInternal *internal; } AVCUDADeviceContext;
int ff_cuda_check_stub(void *avctx, void *cuda_dl, CUresult err, const char *what);
#define FF_CUDA_CHECK_DL(avclass, cudl, x) ff_cuda_check_stub(avclass, cudl, (x), #x)

/* opt.h */
enum AVOptionType { AV_OPT_TYPE_INT, AV_OPT_TYPE_FLOAT, AV_OPT_TYPE_STRING, AV_OPT_TYPE_CONST };
#define AV_OPT_FLAG_FILTERING_PARAM (1 << 16)
#define AV_OPT_FLAG_VIDEO_PARAM 16
typedef struct AVOption {
    const char *name, *help; int offset; enum AVOptionType type;
    union { int64_t i64; double dbl; const char *str; } default_val;
    double min, max; int flags; const char *unit;
} AVOption;
typedef struct AVClass { const char *class_name; const char *(*item_name)(void *); const AVOption *option; int version; } AVClass;
const char *av_default_item_name(void *ctx);
#define AVFILTER_DEFINE_CLASS(fname) static const AVClass fname##_class = { #fname, av_default_item_name, fname##_options, 0 }

/* avfilter.h / internal.h / video.h */
typedef struct AVFilterContext AVFilterContext;
typedef struct AVFilterLink {
    AVFilterContext *src, *dst; int w, h, format; AVBufferRef *hw_frames_ctx;
} AVFilterLink;
struct AVFilterContext { const AVClass *av_class; void *priv; AVFilterLink **inputs, **outputs; };
typedef struct AVFilterPad {
    const char *name; enum AVMediaType type;
    int (*filter_frame)(AVFilterLink *link, AVFrame *frame);
    int (*config_props)(AVFilterLink *link);
} AVFilterPad;
typedef struct AVFilter {
    const char *name, *description; const AVFilterPad *inputs, *outputs; int nb_inputs, nb_outputs;
    const AVClass *priv_class; int flags; int (*init)(AVFilterContext *); void (*uninit)(AVFilterContext *);
    int pix_fmt; int priv_size, flags_internal;
} AVFilter;
#define FILTER_INPUTS(a) .inputs = (a), .nb_inputs = sizeof(a) / sizeof((a)[0])
#define FILTER_OUTPUTS(a) .outputs = (a), .nb_outputs = sizeof(a) / sizeof((a)[0])
#define FILTER_SINGLE_PIXFMT(f) .pix_fmt = (f)
#define FF_FILTER_FLAG_HWFRAME_AWARE 1
int ff_filter_frame(AVFilterLink *link, AVFrame *frame);
#endif
