/*
 * raisr_cuda.h -- the thin C ABI of the B200 RAISR engine (plain pointers and sizes, no C++/torch types).
 *
 * This is what a foreign-function binding of the reference's hot path would bind: RNLHandler_* in
 * raisr/RaisrHandler.h forwards to it, and callers that already hold frames in GPU memory (NVDEC/NVENC
 * pipelines, the bench) use the *_device entry points directly.  Each function cites the reference
 * code it replaces (paths relative to /root/reference/Library).
 *
 * All functions return 0 (RNLErrorNone) or an RNLERRORTYPE value; human-readable diagnostics go to
 * stdout as "[RAISR ERROR] ..." like the reference's.  There is no CPU fallback: without a CUDA device
 * raisr_cuda_create fails.
 */
#ifndef RAISR_CUDA_H
#define RAISR_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct raisr_cuda_engine raisr_cuda_engine;

/* numerics of the bucket hash's three square roots and three divisions */
enum {
    RAISR_NUMERICS_IEEE = 0,   /* sqrt.rn / div.rn: the specification in oracle/raisr_oracle.c (ORACLE_SQRT_IEEE)          */
    RAISR_NUMERICS_X86  = 1,   /* reproduces the compiled reference: vrcp14ps(vrsqrt14ps(x)) etc. via lookup tables          */
    RAISR_NUMERICS_X86_IF_AVAILABLE = 2, /* X86 when the tables are linked in, else IEEE (what RNLHandler_Init asks for)     */
    RAISR_NUMERICS_FP16_FILTER = 3,      /* OPT-IN fast numerics: hash as in mode 2 (buckets identical to the fp32 path), the 121-tap
                                            filter in half precision (coefficients rounded to binary16, HMUL2/HFMA2 chains, fp32 tree);
                                            the counterpart of the reference's asm=avx512fp16 (Raisr_AVX512FP16.cpp:227-242).  8/10 bit
                                            only.  Y is NOT bit-identical to the fp32 path: see DESIGN.md for the measured error.       */
    RAISR_NUMERICS_FAST_HASH = 4         /* OPT-IN experiment: the 11x11 Gaussian structure tensor as two 11-tap passes (the reference's
                                            table is rank one up to 2e-6); hash numerics as in mode 2 behind it.  Buckets are NOT
                                            bit-identical (measured agreement in DESIGN.md); filter and blend are the exact ones.        */
};

/* special values of raisr_cuda_config.device */
enum {
    RAISR_CUDA_DEVICE_CURRENT = -1,         /* the calling thread's current device, re-selected on every call            */
    RAISR_CUDA_DEVICE_CALLER_CONTEXT = -2   /* driver-API interop (FFmpeg's AVCUDADeviceContext): the engine runs in the
                                               CUcontext the caller has made current around EVERY call and never switches  */
};

typedef struct raisr_cuda_config {
    const char *model_path;    /* filter folder: filterbin_2_<bits>[_2], Qfactor_{str,coh}bin_2_<bits>[_2], config            */
    float ratio;               /* 2.0 or 1.5                                                                                  */
    unsigned bit_depth;        /* 8, 10 or 16                                                                                 */
    int range_type;            /* 1 = video range, 2 = full range (RangeType)                                                 */
    unsigned passes;           /* 1 or 2                                                                                      */
    unsigned two_pass_mode;    /* 1: upscale in pass 1; 2: upscale in pass 2                                                  */
    int device;                /* CUDA device ordinal, or RAISR_CUDA_DEVICE_CURRENT / RAISR_CUDA_DEVICE_CALLER_CONTEXT          */
    int numerics;              /* RAISR_NUMERICS_*                                                                            */
    int keep_hash;             /* != 0: keep the per-pass bucket planes for raisr_cuda_read_hash (parity tests)               */
} raisr_cuda_config;

/* Model loading + engine construction.  Replaces RNLInit (Raisr.cpp:1409-1679) and ReadTrainedData
 * (Raisr.cpp:246-433): same files, same validation, same messages. */
int raisr_cuda_create(const raisr_cuda_config *cfg, raisr_cuda_engine **out);

/* Frame geometry.  Replaces RNLSetRes (Raisr.cpp:1681-1829): no row bands, one device plane per image. */
int raisr_cuda_set_res(raisr_cuda_engine *e, unsigned in_w, unsigned in_h, unsigned out_w, unsigned out_h,
                       unsigned in_cw, unsigned in_ch, unsigned out_cw, unsigned out_ch);

/* One frame with HOST planes (steps in bytes), blocking.  Replaces RNLProcess (Raisr.cpp:1294-1397) ->
 * processSegment (Raisr.cpp:890-1289) + the two chroma resizes (Raisr.cpp:1373-1388).  in_u/in_v/out_u/out_v
 * may all be NULL for luma-only use. */
int raisr_cuda_process_host(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, const void *in_u,
                            size_t in_u_step, const void *in_v, size_t in_v_step, void *out_y, size_t out_y_step,
                            void *out_u, size_t out_u_step, void *out_v, size_t out_v_step, int blending);

/* One frame with DEVICE planes, asynchronous on `stream` (a cudaStream_t, NULL = legacy default stream).
 * Same computation as above without the copies. */
int raisr_cuda_process_device(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, const void *in_u,
                              size_t in_u_step, const void *in_v, size_t in_v_step, void *out_y, size_t out_y_step,
                              void *out_u, size_t out_u_step, void *out_v, size_t out_v_step, int blending,
                              void *stream);

/* One frame with DEVICE planes in SEMI-PLANAR layout -- NV12 (8 bit) and P010 (16-bit words, 10-bit value in the high bits:
 * sample_shift = 6), what NVDEC produces and NVENC consumes, and two of the three formats of the reference's hardware-frame filter
 * (vf_raisr_opencl.c:166-169; its kernels apply VideoDataType.bitShift the same way, Raisr.cpp:1313-1348).  in_uv / out_uv hold U and V
 * interleaved; chroma geometry as given to raisr_cuda_set_res (pairs per row = chroma width).  Samples are read as word >> sample_shift
 * and written as value << sample_shift, luma and chroma alike.  Same single launch per frame as the planar entry. */
int raisr_cuda_process_device_semiplanar(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, const void *in_uv,
                                         size_t in_uv_step, void *out_y, size_t out_y_step, void *out_uv, size_t out_uv_step,
                                         int sample_shift, int blending, void *stream);

/* Row-band form of the luma path for multi-GPU sharding (the reference's per-thread bands, Raisr.cpp:1738-1779):
 * computes output rows [row0, row1) only; in_y still points at row 0 of the full input plane, of which only
 * the rows the band depends on are read.  With two passes the first pass is recomputed on the rows the second can
 * reach (overlap-recompute), so bands need no halo exchange. */
int raisr_cuda_process_device_rows(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, void *out_y,
                                   size_t out_y_step, int blending, unsigned row0, unsigned row1, void *stream);

/* Peer-store row bands: a rank may pass, as out_y of raisr_cuda_process_device_rows, a pointer into ANOTHER GPU's frame buffer
 * (same node, NVLink/NVSwitch peer memory).  Its band is then written by the pass kernel's own stores straight into the
 * gathering GPU's memory, tile by tile while the band is being computed -- no collective, no copy after the kernel.  With one
 * process per GPU the buffer crosses the process boundary as a CUDA IPC handle: the owner exports it, the peers open it.
 * (The reference's bands live in one address space, Raisr.cpp:1738-1779; this is the multi-GPU form of "every band writes its
 * rows of the one output frame".)  handle = 64 opaque bytes (cudaIpcMemHandle_t).
 * An IPC handle names a whole allocation: *offset = distance of device_ptr from the allocation's base (sub-allocators such as
 * PyTorch's hand out interior pointers); raisr_cuda_ipc_open returns base + offset in the opening process. */
int raisr_cuda_ipc_export(void *device_ptr, unsigned char handle[64], size_t *offset);
int raisr_cuda_ipc_open(const unsigned char handle[64], size_t offset, void **device_ptr);
int raisr_cuda_ipc_close(void *device_ptr);

/* Bucket plane (int32, -1 where a pixel is not hashed) of pass 0 or 1 of the last frame -> host memory,
 * w*h entries with w,h the plane that pass ran on.  Needs cfg.keep_hash. */
int raisr_cuda_read_hash(raisr_cuda_engine *e, int pass, int32_t *host_out, size_t count);

/* RAISR_NUMERICS_* actually in effect (resolves RAISR_NUMERICS_X86_IF_AVAILABLE) */
int raisr_cuda_numerics(const raisr_cuda_engine *e);

/* number of kernels the engine launched since creation (bench's gpu_launches) */
unsigned long long raisr_cuda_launch_count(const raisr_cuda_engine *e);

/* Replaces RNLDeinit (Raisr.cpp:1842-1909). */
void raisr_cuda_destroy(raisr_cuda_engine *e);

const char *raisr_cuda_version(void);

#ifdef __cplusplus
}
#endif
#endif
