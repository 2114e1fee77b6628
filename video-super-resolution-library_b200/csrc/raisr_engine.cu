// raisr_engine.cu -- host side of the B200 RAISR engine behind include/raisr_cuda.h.
//
// Owns the model tables, the device planes and the launch plan of a frame:
//   passes=1          : [upscale+filter] in -> out
//   passes=2, mode=1  : [upscale+filter, set 1] in -> mid(HR)   ; [filter, set 2] mid -> out
//   passes=2, mode=2  : [filter, set 1] in -> mid(LR)           ; [upscale+filter, set 2] mid -> out
// (the reference's pass control, Library/Raisr.cpp:896-975), plus one resize launch per chroma plane
// (Raisr.cpp:1373-1388).  The pass-1 -> pass-2 dependency is stream order; the intermediate plane is
// quantised to u8/u16 exactly like gIntermediateY (Raisr.cpp:919-927, 1716-1723).
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "raisr/RaisrDefaults.h"
#include "raisr_cuda.h"
#include <cuda_fp16.h>

#include "raisr_kernels.cuh"
#include "raisr_launch.h"
#include "raisr_hostcopy.h"
#include "raisr_model.h"
#include "x86_tables.h"

namespace raisr {

#define CUDA_OK(call)                                                                                          \
    do {                                                                                                       \
        cudaError_t err__ = (call);                                                                            \
        if (err__ != cudaSuccess) {                                                                            \
            std::cout << "[RAISR ERROR] CUDA failure: " << cudaGetErrorString(err__) << " at " << __FILE__ << ":" \
                      << __LINE__ << std::endl;                                                                \
            return RNLErrorInsufficientResources;                                                              \
        }                                                                                                      \
    } while (0)

// One axis of the cheap upscale: dst index d samples src at ((2d+1)*src - dst) / (2*dst), replicate border,
// weights kept as exact integers over a common (gcd-reduced) denominator.  Definition: oracle/ipp_standin/ipp.h.
struct AxisMap {
    std::vector<int> map, w;   // map[d] = i0 << 1 | (i1 != i0);  w[d] = weight numerator of tap i1
    int den = 1;
    int *d_map = nullptr, *d_w = nullptr;
    void build(int src, int dst)
    {
        map.resize(dst);
        w.resize(dst);
        const long long D = 2LL * dst;
        long long g = D;
        for (int d = 0; d < dst; ++d) {
            const long long num = (2LL * d + 1) * src - dst;
            const long long q = num >= 0 ? num / D : -((-num + D - 1) / D);
            const long long r = num - q * D;
            long long a = q, b = q + 1;
            a = std::min<long long>(std::max<long long>(a, 0), src - 1);
            b = std::min<long long>(std::max<long long>(b, 0), src - 1);
            map[d] = (int)(a << 1) | (b != a);
            w[d] = (int)r;
            g = std::gcd(g, r);
        }
        den = (int)(D / g);
        for (int d = 0; d < dst; ++d) w[d] = (int)(w[d] / g);
    }
    int upload()
    {
        release();
        CUDA_OK(cudaMalloc(&d_map, map.size() * sizeof(int)));
        CUDA_OK(cudaMalloc(&d_w, w.size() * sizeof(int)));
        CUDA_OK(cudaMemcpy(d_map, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d_w, w.data(), w.size() * sizeof(int), cudaMemcpyHostToDevice));
        return 0;
    }
    void release()
    {
        cudaFree(d_map);
        cudaFree(d_w);
        d_map = d_w = nullptr;
    }
};

struct Plane {
    void *ptr = nullptr;
    size_t pitch = 0;
    int w = 0, h = 0;
    int alloc(int w_, int h_, int bps)
    {
        release();
        w = w_; h = h_;
        CUDA_OK(cudaMallocPitch(&ptr, &pitch, (size_t)w * bps, h));
        return 0;
    }
    void release() { cudaFree(ptr); ptr = nullptr; }
};

// hashed column range of a row: the reference's column loop (Raisr.cpp:1065-1066,1246-1250) run symbolically
static void hashed_cols(int W, int *c_end, int *tail_start, int *ov_end)
{
    int step = 16, c = 6, tail = -1, last16 = -1;
    while (c + step <= W - 6) {
        if (step == 16) last16 = c; else if (tail < 0) tail = c;
        if (step > 8 && c + 32 > W - 6) step = 8;
        c += step;
    }
    if (tail < 0) tail = c;
    *c_end = c;
    *tail_start = tail;
    *ov_end = last16 >= 0 ? std::min(last16 + 16, c) : tail;
}

// Copy threads of the pageable-plane path: FFmpeg's software frames are ordinary malloc'ed memory, which the copy engines cannot
// reach.  Instead of the driver's serial bounce-buffer copies (measured: 333 frames/s at 1080p->4K against 1520 with page-locked
// planes) the engine DMAs from/to page-locked staging planes of its own and moves the bytes between those and the caller's planes
// with a few host threads, band by band, while the kernel is still running.
class CopyPool {
public:
    void start(int n, int device, bool bind)
    {
        if (const char *z = std::getenv("RAISR_CUDA_COPY_POLL_US")) poll_us_ = std::max(0, std::min(100000, std::atoi(z)));
        for (int i = 0; i < n; ++i)
            threads_.emplace_back([this, device, bind] {
                if (bind) cudaSetDevice(device);
                for (;;) {
                    Job job;
                    // Inside a frame the next job is at most a band (~30 us) away: poll for that long before sleeping, so that
                    // neither the submitting thread pays a futex wake-up per job nor the job its latency (the rows delivered
                    // after the kernel's end and the rows staged ahead of its launch are on the frame's critical path).
                    for (auto t0 = std::chrono::steady_clock::now(); queued_.load(std::memory_order_acquire) == 0;) {
                        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(poll_us_)) break;
                        __builtin_ia32_pause();
                    }
                    {
                        std::unique_lock<std::mutex> lk(m_);
                        cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                        if (q_.empty()) return;
                        queued_.fetch_sub(1, std::memory_order_relaxed);
                        // the oldest job whose event has completed; if none has, the oldest one (block on its event).  A thread
                        // never sleeps on a late event (the chroma planes' D2H) while an earlier band is ready to be delivered.
                        auto pick = q_.begin();
                        for (auto it = q_.begin(); it != q_.end(); ++it)
                            if (!it->ev || cudaEventQuery(it->ev) != cudaErrorNotReady) { pick = it; break; }
                        job = std::move(*pick);
                        q_.erase(pick);
                    }
                    if (job.ev) cudaEventSynchronize(job.ev);
                    job.fn();
                    {
                        std::lock_guard<std::mutex> lk(m_);
                        --pending_;
                        if (--group_pending_[job.group] == 0) done_.notify_all();
                    }
                }
            });
    }
    // fn runs on a pool thread once ev (may be null) has completed; jobs of one group (0 .. kGroups-1) can be waited for on their own
    void submit(cudaEvent_t ev, std::function<void()> fn, int group = 0)
    {
        if (threads_.empty()) {
            if (ev) cudaEventSynchronize(ev);
            fn();
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(Job{ev, std::move(fn), group});
            queued_.fetch_add(1, std::memory_order_release);
            ++pending_;
            ++group_pending_[group];
        }
        cv_.notify_one();
    }
    // (the waits poll for a moment before sleeping, like the workers: the caller's thread is on the frame's critical path)
    void wait_group(int group)
    {
        if (poll([this, group] { return group_pending_[group] == 0; })) return;
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this, group] { return group_pending_[group] == 0; });
    }
    void wait_all()
    {
        if (poll([this] { return pending_ == 0; })) return;
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }

private:
    template <typename Pred> bool poll(Pred done)
    {
        for (auto t0 = std::chrono::steady_clock::now(); std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(2 * poll_us_);) {
            {
                std::lock_guard<std::mutex> lk(m_);
                if (done()) return true;
            }
            for (int i = 0; i < 32; ++i) __builtin_ia32_pause();
        }
        return false;
    }
    static constexpr int kGroups = 3;
    struct Job {
        cudaEvent_t ev = nullptr;
        std::function<void()> fn;
        int group = 0;
    };
    int group_pending_[kGroups] = {};
    std::atomic<int> queued_{0};                // jobs in q_ (polled without the lock)
    int poll_us_ = 100;                         // RAISR_CUDA_COPY_POLL_US: how long an idle worker polls before it sleeps (0: sleep at once)
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::deque<Job> q_;
    int pending_ = 0;
    bool stop_ = false;
};

// (raisr_hostcopy.cpp: streaming stores -- every destination of this path is written once and read by someone else)
static void copy_rows(void *dst, size_t dstep, const void *src, size_t sstep, size_t row_bytes, int rows)
{
    host_copy_rows(dst, dstep, src, sstep, row_bytes, rows);
}

}  // namespace raisr

using namespace raisr;

struct raisr_cuda_engine {
    raisr_cuda_config cfg{};
    std::string model_path;
    Model model;
    int bps = 1;                    // bytes per sample
    int lo = 0, hi = 255;
    int device = 0;
    bool bind_device = true;        // false (cfg.device == RAISR_CUDA_DEVICE_CALLER_CONTEXT): run in whatever CUDA context the caller made current, never switch
    int num_sms = 148;
    int blending = 2;               // BlendingMode of the frame being processed
    int sample_shift = 0;           // frame being processed: samples carry their value in the high bits (P010: 6); 0 for everything else
    bool no_memops = false;         // RAISR_CUDA_NO_MEMOPS=1 (or CUDA_LAUNCH_BLOCKING=1): no in-kernel flag waits, everything in plain stream order
    bool split_h2d = true;          // pipelined kernel: input plane in two copies, the second one under the kernel; RAISR_CUDA_SPLIT_H2D=0 disables
    bool tail_in_place = true;      // rows of the last round of tiles go straight into a page-locked output plane (no copy after the kernel); RAISR_CUDA_TAIL_IN_PLACE=0 disables
    bool band_pipeline = true;      // luma D2H in row bands under the kernel; RAISR_CUDA_NO_BAND_PIPELINE=1 disables
    bool use_pipe = true;           // persistent warp-specialised kernel; RAISR_CUDA_KERNEL=tile selects the phase-sequential kernel (cross-check)
    float gw[11][6] = {};           // folded Gaussian weights of this engine's bit depth (copied into every launch's parameters)
    bool test_drop_in_flag = false; // RAISR_CUDA_TEST_DROP_IN_FLAG=1 (tests only): drop the split-H2D flag of the first frame
    bool fast_hash = false;         // RAISR_NUMERICS_FAST_HASH: separable structure tensor (opt-in, buckets not bit-identical)
    bool filter_fp16 = false;       // RAISR_NUMERICS_FP16_FILTER: half-precision filter stage (opt-in, Y not bit-identical); the hash keeps cfg.numerics
    int chain_passes = -1;          // two-pass configurations in one persistent launch: -1 = where measured faster (see launch_chained),
                                    // RAISR_CUDA_CHAIN=1: always, RAISR_CUDA_CHAIN=0: never (one launch per pass)
    bool coop_launch = false;       // device supports cooperative launches (needed by the chained launch)
    unsigned *d_rows_done = nullptr; int rows_done_cap = 0;   // chained launch: finished tiles per tile row of the first pass
    void *d_filters16[2] = {nullptr, nullptr};                 // fp16 copies of the filter tables (filter_fp16 only)
    float *d_filters[2] = {nullptr, nullptr};
    void *d_lut[4] = {nullptr, nullptr, nullptr, nullptr};   // rsqrt14 runs, rcp14 runs, rsqrtps, rcpps
    // geometry
    bool have_res = false;
    int in_w = 0, in_h = 0, out_w = 0, out_h = 0, in_cw = 0, in_ch = 0, out_cw = 0, out_ch = 0;
    AxisMap yx, yy, cx, cy;
    int up_src_h = 0;
    float quarter = 0.25f, half = 0.5f;
    Plane d_in[3], d_out[3], d_mid;
    int *d_hash[2] = {nullptr, nullptr};
    int hash_w[2] = {0, 0}, hash_h[2] = {0, 0};
    cudaStream_t stream = nullptr, stream_uv = nullptr;
    cudaEvent_t ev_uv = nullptr, ev_in = nullptr;
    unsigned long long launches = 0;
    // host-pointer pipeline: the final pass signals finished row bands, the D2H stream waits on the counters
    static constexpr int kMaxBands = 16;
    unsigned *d_band_done = nullptr;
    cudaStream_t stream_d2h = nullptr;
    typedef int (*WaitValue32Fn)(cudaStream_t, unsigned long long, unsigned, unsigned);
    WaitValue32Fn wait_value32 = nullptr, write_value32 = nullptr;
    // cuTensorMapEncodeTiled (driver entry point, like the stream memory operations): tensor maps of the passes' input planes for the
    // TMA load of stage A; the last few encodings are cached by (pointer, pitch, geometry)
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
    EncodeTiledFn encode_tiled = nullptr;
    bool use_tma = true;            // RAISR_CUDA_TMA=0: clamped scalar loads everywhere
    struct TmapEntry { const void *ptr = nullptr; size_t pitch = 0; int w = 0, h = 0; alignas(64) CUtensorMap map; };
    TmapEntry tmaps[4];
    int tmap_next = 0;
    unsigned *d_in_ready = nullptr;      // watermark of the split H2D: (frame_seq << 16) | input rows that have arrived
    unsigned frame_seq = 0;
    // pageable caller planes: page-locked staging planes + copy threads (RAISR_CUDA_STAGE_PAGEABLE=0 / RAISR_CUDA_COPY_THREADS=n)
    bool stage_pageable = true;
    int copy_threads = 4;
    void *h_stage_in[3] = {nullptr, nullptr, nullptr}, *h_stage_out[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_band[kMaxBands] = {}, ev_chroma_out = nullptr;
    CopyPool *pool = nullptr;
    unsigned *h_err = nullptr, *d_err = nullptr;   // page-locked, device-mapped word: an in-kernel flag wait timed out (checked after every frame)
    // RAISR_CUDA_TIMING=1: per-stage device times of the host-pointer call, printed when the engine is destroyed
    bool timing = false; cudaEvent_t tev[3] = {nullptr, nullptr, nullptr}; double t_h2d = 0, t_kern = 0; unsigned long long t_n = 0;
    double t_host[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // RAISR_CUDA_TIMING: host-side phase marks of process_host_pipe (sums of microseconds since entry)
    unsigned *d_chroma_ready = nullptr;  // [0] sequence number of the last frame whose chroma planes have arrived (H2D on stream_uv), [1] CTAs done with them (running total)
    unsigned chroma_seq = 0, chroma_done_target = 0;
    unsigned band_target[kMaxBands] = {};  // running totals the D2H stream waits for, per output row band
    cudaStream_t stream_h2d = nullptr;
    int last_chroma_ctas = 0;               // CTAs of the most recent launch that carried a chroma job (each bumps chroma_done once)
    // row bands of the most recent launch with a band_done counter: computed ONCE (launch_pass_t) and used both for the kernel's
    // band_tiles_y and for the host's wait targets / copy ranges
    int n_bands = 0;
    int band_row0[kMaxBands] = {}, band_row1[kMaxBands] = {};
    unsigned band_tiles[kMaxBands] = {};
};

namespace {

void fill_weights(raisr_cuda_engine *e, unsigned bits)
{
    // the 21 distinct literals of gGaussian2DOriginal (Raisr_globals.h:213-224), G[i][j] = kG[min][max] after folding
    static const double kG[6][6] = {
        {7.76554e-05, 0.000239195, 0.0005738, 0.001072, 0.00155975, 0.00176743},
        {0, 0.000736774, 0.00176743, 0.00330199, 0.00480437, 0.00544406},
        {0, 0, 0.00423984, 0.00792107, 0.0115251, 0.0130596},
        {0, 0, 0, 0.0147985, 0.0215317, 0.0243986},
        {0, 0, 0, 0, 0.0313284, 0.0354998},
        {0, 0, 0, 0, 0, 0.0402265}};
    const float M = bits == 8 ? 255.0f : bits == 10 ? 1023.0f : 65535.0f;
    const float NF = 1.0f / (M * M * 2.0f * 2.0f);                       // NF_8 / NF_10 / NF_16, Raisr_globals.h:204-206
    for (int i = 0; i < 11; ++i) {
        const int ii = i < 5 ? i : 10 - i;
        for (int m = 0; m < 6; ++m) e->gw[i][m] = (float)((double)NF * kG[std::min(ii, m)][std::max(ii, m)]);
    }
}

// ---- launch planning ------------------------------------------------------------------------------------------------
// One pass, planned: tile height, grid, band geometry, upscale flavour.  The kernels live in their own translation units
// (raisr_launch.h); this file only fills parameter blocks.
struct PassPlan {
    PassParams q;
    dim3 grid;               // tiles (x) by tile rows (y)
    int ups = 0;             // 0 none, 1 exact 2x, 2 axis maps
};

// Tensor map of an input plane for the pipelined kernel's stage A: 2-D, one element per sample, box = the low-res window of one
// tile (tmap_box_w x TMAP_BOX_H), out-of-range elements zero (such tiles take the scalar path anyway).  nullptr when the plane does
// not meet TMA's alignment rules (16-byte base and pitch) or the driver refuses: the kernel then keeps its scalar loads.
const CUtensorMap *input_tensor_map(raisr_cuda_engine *e, const void *ptr, size_t pitch, int w, int h)
{
    static_assert(sizeof(CUtensorMap) == sizeof(PassParams::in_tmap), "tensor map size");
    if ((reinterpret_cast<uintptr_t>(ptr) % 16) != 0 || (pitch % 16) != 0 || w < 1 || h < 1) return nullptr;
    for (auto &t : e->tmaps)
        if (t.ptr == ptr && t.pitch == pitch && t.w == w && t.h == h) return &t.map;
    raisr_cuda_engine::TmapEntry &t = e->tmaps[e->tmap_next];
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}, strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)(e->bps == 1 ? TmapBox<uint8_t>::W : TmapBox<uint16_t>::W), (cuuint32_t)TMAP_BOX_H}, estr[2] = {1, 1};
    const CUresult rc = e->encode_tiled(&t.map, e->bps == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(ptr), dims,
                                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { t.ptr = nullptr; return nullptr; }
    t.ptr = ptr; t.pitch = pitch; t.w = w; t.h = h;
    e->tmap_next = (e->tmap_next + 1) % 4;
    return &t.map;
}

PassPlan plan_pass(raisr_cuda_engine *e, const PassParams &p)
{
    // tile height: the largest even th <= the kernel's maximum whose tile count still fits the same number of waves (one CTA per SM)
    PassPlan pl;
    PassParams &q = pl.q;
    q = p;
    const int bpsx = e->bps;
    const int rows = p.row1 - p.row0, gx = (p.W + TW - 1) / TW;
    const int thmax = e->use_pipe ? pipe_tile_h_max() : TH_MAX;
    int ny = (rows + thmax - 1) / thmax;
    const int waves = (gx * ny + e->num_sms - 1) / e->num_sms;
    while ((long long)gx * (ny + 1) <= (long long)waves * e->num_sms && 2 * (ny + 1) <= rows) ++ny;
    q.tile_h = std::min(thmax, (((rows + ny - 1) / ny) + 1) & ~1);
    pl.grid = dim3(gx, (rows + q.tile_h - 1) / q.tile_h);
    // tile rows of the last round of the persistent kernel (they finish together, at the very end): written in place into the
    // caller's pinned plane when there is one, so that no copy remains after the kernel; the bands cover the rows above
    int banded_rows = (int)pl.grid.y;
    if (q.out_tail && e->use_pipe) {
        banded_rows = std::max(0, (int)pl.grid.y - ((e->num_sms + gx - 1) / gx + 1));
        q.tail_row0 = p.row0 + banded_rows * q.tile_h;
    } else {
        q.out_tail = nullptr;
    }
    if (q.band_done) {
        // the ONE place the band geometry is decided: tile rows per band for the kernel, rows and tile counts per band for the host
        constexpr int kMax = raisr_cuda_engine::kMaxBands;
        q.band_tiles_y = std::max(1, (banded_rows + kMax - 1) / kMax);
        e->n_bands = 0;
        for (int ty = 0; ty < banded_rows; ty += q.band_tiles_y) {
            const int tiles_y = std::min(q.band_tiles_y, banded_rows - ty), b = e->n_bands++;
            e->band_row0[b] = p.row0 + ty * q.tile_h;
            e->band_row1[b] = std::min(p.row1, p.row0 + (ty + tiles_y) * q.tile_h);
            e->band_tiles[b] = (unsigned)(tiles_y * gx);
        }
    }
    q.vec_store = ((reinterpret_cast<uintptr_t>(p.out) | p.out_pitch | reinterpret_cast<uintptr_t>(q.out_tail) | (q.out_tail ? q.out_tail_pitch : 0)) % (4 * bpsx)) == 0;
    // 2x fast path: exact factor 2 in both axes and even band origin
    const bool fast2x = p.upscale && p.W == 2 * p.in_w && p.denx == 4 && p.deny == 4 && (p.row0 % 2) == 0 && p.up_src_h * 2 == p.H;
    pl.ups = !p.upscale ? 0 : (fast2x ? 1 : 2);
    q.use_tmap = 0;
    if (e->use_pipe && pl.ups == 1 && e->use_tma && e->encode_tiled) {
        if (const CUtensorMap *m = input_tensor_map(e, p.in, p.in_pitch, p.in_w, p.in_h)) {
            memcpy(q.in_tmap, m, sizeof(CUtensorMap));
            q.use_tmap = 1;
        }
    }
    return pl;
}

int cuda_failed(int err, const char *what)
{
    std::cout << "[RAISR ERROR] CUDA failure: " << cudaGetErrorString((cudaError_t)err) << " (" << what << ")" << std::endl;
    return RNLErrorInsufficientResources;
}

// instantiation of the pipelined kernel (raisr_pipe_kernel.cuh "NV"): 1 fp16 filter stage, 2 separable fast hash, 4 exact with the X86
// numerics compiled in, 0 exact with the numerics decided at run time (IEEE)
int pipe_variant(const raisr_cuda_engine *e)
{
    return e->filter_fp16 ? 1 : e->fast_hash ? 2 : e->cfg.numerics == RAISR_NUMERICS_X86 ? 4 : 0;
}

int launch_frame(raisr_cuda_engine *e, const FrameLaunch &fl)
{
    const int nv = pipe_variant(e);
    if (e->bps == 1)
        return nv == 1 ? launch_frame_pipe<uint8_t, 1>(fl) : nv == 2 ? launch_frame_pipe<uint8_t, 2>(fl) : nv == 4 ? launch_frame_pipe<uint8_t, 4>(fl) : launch_frame_pipe<uint8_t, 0>(fl);
    return nv == 1 ? launch_frame_pipe<uint16_t, 1>(fl) : nv == 2 ? launch_frame_pipe<uint16_t, 2>(fl) : nv == 4 ? launch_frame_pipe<uint16_t, 4>(fl) : launch_frame_pipe<uint16_t, 0>(fl);
}

// one pass = one launch
int launch_pass(raisr_cuda_engine *e, const PassParams &p, cudaStream_t s)
{
    PassPlan pl = plan_pass(e, p);
    int err;
    if (e->use_pipe) {
        // persistent warp-specialised kernel: one CTA per SM walks the tiles (producer/consumer warp groups)
        FrameLaunch fl;
        fl.a = pl.q; fl.ups_a = pl.ups; fl.stream = s;
        fl.grid = std::min((int)(pl.grid.x * pl.grid.y), e->num_sms);
        if (pl.q.chroma_n) e->last_chroma_ctas = fl.grid;
        err = launch_frame(e, fl);
    } else {
        err = e->bps == 1 ? launch_pass_tile<uint8_t>(pl.q, pl.ups, pl.grid, s) : launch_pass_tile<uint16_t>(pl.q, pl.ups, pl.grid, s);
    }
    if (err) return cuda_failed(err, "pass launch");
    e->launches++;
    return 0;
}

// Both passes of a two-pass configuration in ONE persistent launch (Raisr.cpp:896-927 runs them back to back per band, with
// a neighbour wait, :905-916): the CTAs walk pass 1's tiles and carry straight on into pass 2's; a pass-2 tile waits for the
// pass-1 tile rows that cover its input rows (counters in d_rows_done).  The intermediate plane stays in L2.
// Returns -1 when this pair has no chained instantiation or the device cannot launch it cooperatively (-> two launches).
int launch_chained(raisr_cuda_engine *e, const PassParams &p1, const PassParams &p2, cudaStream_t s)
{
    if (!e->use_pipe || !e->chain_passes || !e->coop_launch || !e->d_rows_done) return -1;
    PassPlan a = plan_pass(e, p1), b = plan_pass(e, p2);
    if ((int)a.grid.y > e->rows_done_cap) return -1;
    // Default policy = what was measured (same box, luma passes of a frame, chained vs one launch per pass): mode 2 at 2x
    // 0.732 vs 0.761 ms (1080p->4K) and 2.636 vs 2.670 ms (4K->8K), 1.5x 0.285 vs 0.296 ms -- chained wins; mode 1 at 2x (exact-2x
    // upscale in pass 1, plain pass 2) 1.173 vs 1.151 ms -- the boundary-free launch loses 2 % to the register allocation of its
    // second inlined pass, so that pair runs as two launches unless RAISR_CUDA_CHAIN=1 asks for the chained form.
    if (e->chain_passes < 0 && p1.ptypes == 4 && a.ups == 1 && b.ups == 0) return -1;
    FrameLaunch fl;
    fl.a = a.q; fl.b = b.q; fl.two = true; fl.ups_a = a.ups; fl.ups_b = b.ups; fl.stream = s;
    fl.a.rows_done = e->d_rows_done;
    fl.b.dep_done = e->d_rows_done;
    fl.b.dep_gx = (int)a.grid.x; fl.b.dep_ny = (int)a.grid.y;
    fl.b.dep_row0 = a.q.row0; fl.b.dep_row1 = a.q.row1; fl.b.dep_th = a.q.tile_h;
    fl.grid = std::min((int)(a.grid.x * a.grid.y + b.grid.x * b.grid.y), e->num_sms);
    if (fl.a.chroma_n) e->last_chroma_ctas = fl.grid;          // every CTA of the grid takes a share of the chroma planes (and bumps chroma_done once)
    if (cudaMemsetAsync(e->d_rows_done, 0, sizeof(unsigned) * a.grid.y, s) != cudaSuccess) return cuda_failed((int)cudaGetLastError(), "rows_done memset");
    const int err = launch_frame(e, fl);
    if (err == (int)cudaErrorInvalidDeviceFunction || err == (int)cudaErrorCooperativeLaunchTooLarge) { cudaGetLastError(); return -1; }
    if (err) return cuda_failed(err, "chained launch");
    e->launches++;
    return 0;
}

int launch_resize(raisr_cuda_engine *e, const void *in, size_t in_pitch, void *out, size_t out_pitch, cudaStream_t s)
{
    ResizeParams rp{in, in_pitch, e->in_cw, e->in_ch, out, out_pitch, e->out_cw, e->out_ch,
                    e->cx.d_map, e->cx.d_w, e->cy.d_map, e->cy.d_w, e->cx.den, e->cy.den};
    const dim3 grid((rp.W + 63) / 64, (rp.H + 15) / 16);
    if (e->bps == 1) resize_kernel<uint8_t><<<grid, 256, 0, s>>>(rp);
    else resize_kernel<uint16_t><<<grid, 256, 0, s>>>(rp);
    CUDA_OK(cudaGetLastError());
    e->launches++;
    return 0;
}

// fills the per-pass constant part of PassParams
void pass_common(const raisr_cuda_engine *e, int pass_idx, int W, PassParams *p)
{
    const PassModel &pm = e->model.pass[pass_idx];
    p->filters = e->filter_fp16 ? static_cast<const float *>(e->d_filters16[pass_idx]) : e->d_filters[pass_idx];
    p->ptypes = pm.ptypes;
    p->nbuckets = pm.buckets;
    p->qstr0 = pm.qstr[0]; p->qstr1 = pm.qstr[1];
    p->qcoh0 = pm.qcoh[0]; p->qcoh1 = pm.qcoh[1];
    p->lo = e->lo; p->hi = e->hi;
    hashed_cols(W, &p->c_end, &p->tail_start, &p->ov_end);
    p->numerics = e->cfg.numerics;
    if (e->cfg.numerics == RAISR_NUMERICS_X86) {
        // as compiled: gQAngle = angles * (1/PI); the 8-wide hash divides by 4 and 2 through rcpps + one Newton step
        const float rpi = 1.0f / 3.141592653f;
        p->qangle = (float)e->model.q_angle * rpi;
        p->quarter = e->quarter;
        p->half = e->half;
    } else {
        p->qangle = (float)e->model.q_angle / 3.141592653f;  // gQAngle = gQuantizationAngle / PI (Raisr.cpp:1553)
    }
    p->nangles = e->model.q_angle;
    p->hash_out = e->d_hash[pass_idx];
    p->blending = e->blending;
    p->lut_rsqrt14 = static_cast<const uint2 *>(e->d_lut[0]); p->lut_rcp14 = static_cast<const uint2 *>(e->d_lut[1]);
    p->lut_rsqrtps = static_cast<const uint16_t *>(e->d_lut[2]); p->lut_rcpps = static_cast<const uint16_t *>(e->d_lut[3]);
    memcpy(p->gw, e->gw, sizeof(p->gw));
    p->err_flag = e->d_err;
}

void set_upscale(const raisr_cuda_engine *e, PassParams *p)
{
    p->upscale = 1;
    p->xmap = e->yx.d_map; p->xw = e->yx.d_w; p->ymap = e->yy.d_map; p->yw = e->yy.d_w;
    p->denx = e->yx.den; p->deny = e->yy.den;
    p->up_src_h = e->up_src_h;
}

// chroma planes handed to the pipelined kernel (resized by its producer warps, no launch of their own)
struct ChromaJob {
    const void *in[2]; size_t in_step[2];
    void *out[2]; size_t out_step[2];
    int comps = 1;                             // 2: ONE semi-planar plane in in[0] / out[0] (U and V interleaved)
    const unsigned *ready; unsigned seq;       // optional H2D completion flag
    unsigned *done;                            // optional per-CTA completion counter
};

void set_chroma(const raisr_cuda_engine *e, const ChromaJob *c, PassParams *p)
{
    if (!c) return;
    p->chroma_n = c->comps == 2 ? 1 : 2;
    p->c_comps = c->comps;
    p->c_shift = e->sample_shift;
    for (int i = 0; i < 2; ++i) {
        p->chroma[i].in = c->in[i]; p->chroma[i].in_pitch = c->in_step[i];
        p->chroma[i].out = c->out[i]; p->chroma[i].out_pitch = c->out_step[i];
    }
    p->c_in_w = e->in_cw; p->c_in_h = e->in_ch; p->c_W = e->out_cw; p->c_H = e->out_ch;
    p->c_xmap = e->cx.d_map; p->c_xw = e->cx.d_w; p->c_ymap = e->cy.d_map; p->c_yw = e->cy.d_w;
    p->c_denx = e->cx.den; p->c_deny = e->cy.den;
    p->chroma_ready = c->ready; p->chroma_seq = c->seq; p->chroma_done = c->done;
}

// the luma launch plan; rows [row0,row1) of the final plane (row bands only for single-pass configurations)
int run_luma(raisr_cuda_engine *e, const void *in_y, size_t in_step, void *out_y, size_t out_step, int row0, int row1,
             cudaStream_t s, unsigned *band_done = nullptr, const unsigned *in_ready = nullptr,
             const ChromaJob *chroma = nullptr, void *out_tail = nullptr, size_t out_tail_step = 0, int in_split_row = 0)
{
    const bool two = e->cfg.passes == 2;
    const bool mode2 = two && e->cfg.two_pass_mode == 2;
    if (!two) {
        PassParams p{};
        set_chroma(e, chroma, &p);
        p.in = in_y; p.in_pitch = in_step; p.in_w = e->in_w; p.in_h = e->in_h;
        p.out = out_y; p.out_pitch = out_step; p.W = e->out_w; p.H = e->out_h; p.row0 = row0; p.row1 = row1;
        pass_common(e, 0, p.W, &p);
        set_upscale(e, &p);
        p.in_shift = p.out_shift = e->sample_shift;
        p.band_done = band_done; p.out_tail = out_tail; p.out_tail_pitch = out_tail_step;
        p.in_ready = in_ready; p.in_seq = e->frame_seq; p.in_split_row = in_split_row;
        return launch_pass(e, p, s);
    }
    // Two passes on a row band: pass 1 is recomputed on the rows pass 2 can reach (+-7 output rows of pass 2, mapped back
    // through the upscale in mode 2), so bands stay independent -- overlap-recompute instead of a halo exchange.
    int m0 = 0, m1 = e->d_mid.h;
    if (row0 != 0 || row1 != e->out_h) {
        const int lo = std::max(0, row0 - 7), hi = std::min(e->out_h, row1 + 7);
        if (mode2) {
            m0 = std::max(0, (int)std::floor(lo / e->cfg.ratio) - 2);
            m1 = std::min(e->d_mid.h, (int)std::ceil(hi / e->cfg.ratio) + 2);
        } else {
            m0 = lo & ~1;
            m1 = hi;
        }
    }
    PassParams p1{}, p2{};
    p1.in = in_y; p1.in_pitch = in_step; p1.in_w = e->in_w; p1.in_h = e->in_h;
    p1.out = e->d_mid.ptr; p1.out_pitch = e->d_mid.pitch; p1.W = e->d_mid.w; p1.H = e->d_mid.h; p1.row0 = m0; p1.row1 = m1;
    pass_common(e, 0, p1.W, &p1);
    if (!mode2) set_upscale(e, &p1);
    p2.in = e->d_mid.ptr; p2.in_pitch = e->d_mid.pitch; p2.in_w = e->d_mid.w; p2.in_h = e->d_mid.h;
    p2.out = out_y; p2.out_pitch = out_step; p2.W = e->out_w; p2.H = e->out_h; p2.row0 = row0; p2.row1 = row1;
    pass_common(e, 1, p2.W, &p2);
    if (mode2) set_upscale(e, &p2);
    p1.in_shift = e->sample_shift;               // the intermediate plane holds plain values
    p2.out_shift = e->sample_shift;
    p2.band_done = band_done; p2.out_tail = out_tail; p2.out_tail_pitch = out_tail_step;
    p1.in_ready = in_ready; p1.in_seq = e->frame_seq; p1.in_split_row = in_split_row;
    set_chroma(e, chroma, &p1);                  // resized while pass 1's filter warps finish
    int rc = launch_chained(e, p1, p2, s);       // both passes in one persistent launch where possible
    if (rc >= 0) return rc;
    rc = launch_pass(e, p1, s);
    if (rc) return rc;
    return launch_pass(e, p2, s);
}

// true (and the device alias) when p is page-locked host memory the device can address (cudaHostAlloc / cudaHostRegister)
bool mapped_host_pointer(const void *p, const void **dev)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    *dev = a.devicePointer;
    return true;
}

int check_blending(int blending)
{
    if (blending == CountOfBitsChanged || blending == Randomness) return 0;
    std::cout << "[RAISR ERROR] unknown blending mode " << blending << std::endl;
    return RNLErrorBadParameter;
}

}  // namespace

extern "C" {

const char *raisr_cuda_version(void) { return "raisr-b200 0.1 (API 23.11)"; }

int raisr_cuda_create(const raisr_cuda_config *cfg, raisr_cuda_engine **out)
{
    if (!cfg || !out || !cfg->model_path) return RNLErrorBadParameter;
    *out = nullptr;
    unsigned passes = cfg->passes, mode = cfg->two_pass_mode;
    // pass control, Raisr.cpp:1429-1439
    if (passes == 2) {
        std::cout << "--------------- running 2 pass ---------------\n";
        if (mode != 1 && mode != 2) {
            std::cout << "[RAISR ERROR] Only support two pass mode 1 or 2. " << std::endl;
            return RNLErrorUndefined;
        }
    } else if (passes == 1 && mode == 2) {
        std::cout << "[RAISR WARNING] 1 pass with upscale in 2d pass, mode = 2 ignored !" << std::endl;
        mode = 1;
    } else if (passes != 1) {
        std::cout << "[RAISR ERROR] Only support passes 1 or 2. " << std::endl;
        return RNLErrorUndefined;
    } else {
        mode = 1;
    }
    if (cfg->bit_depth != 8 && cfg->bit_depth != 10 && cfg->bit_depth != 16) {     // Raisr.cpp:1446-1474
        std::cout << "[RAISR ERROR] bit depth: " << cfg->bit_depth << "bits is NOT supported." << std::endl;
        return RNLErrorBadParameter;
    }
    raisr_cuda_engine *e = new raisr_cuda_engine;
    e->cfg = *cfg;
    e->cfg.passes = passes;
    e->cfg.two_pass_mode = mode;
    e->model_path = cfg->model_path;
    e->cfg.model_path = e->model_path.c_str();
    e->bps = cfg->bit_depth == 8 ? 1 : 2;
    const bool video = cfg->range_type == VideoRange;
    if (cfg->bit_depth == 8) { e->lo = video ? 16 : 0; e->hi = video ? 235 : 255; }
    else if (cfg->bit_depth == 10) { e->lo = video ? 64 : 0; e->hi = video ? 940 : 1023; }
    else { e->lo = 0; e->hi = 65535; }

    int rc = load_model(e->model_path, cfg->ratio, cfg->bit_depth, passes, &e->model);
    if (rc != RNLErrorNone) { delete e; return rc; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        std::cout << "[RAISR ERROR] no CUDA device: the B200 engine has no CPU path" << std::endl;
        delete e;
        return RNLErrorUndefined;
    }

    auto fail = [&](int code) { raisr_cuda_destroy(e); return code; };
    e->bind_device = cfg->device != RAISR_CUDA_DEVICE_CALLER_CONTEXT;
    if (cfg->device >= 0) {
        if (cudaSetDevice(cfg->device) != cudaSuccess) {
            std::cout << "[RAISR ERROR] cannot select CUDA device " << cfg->device << std::endl;
            return fail(RNLErrorBadParameter);
        }
    }
    if (cudaGetDevice(&e->device) != cudaSuccess) return fail(RNLErrorInsufficientResources);
    cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->device);
    if (const char *z = std::getenv("RAISR_CUDA_TAIL_IN_PLACE")) e->tail_in_place = std::atoi(z) != 0;
    if (const char *z = std::getenv("RAISR_CUDA_NO_BAND_PIPELINE")) e->band_pipeline = std::atoi(z) == 0;
    if (const char *sp = std::getenv("RAISR_CUDA_SPLIT_H2D")) e->split_h2d = std::atoi(sp) != 0;
    if (const char *nm = std::getenv("RAISR_CUDA_NO_MEMOPS")) e->no_memops = std::atoi(nm) != 0;
    if (const char *lb = std::getenv("CUDA_LAUNCH_BLOCKING")) e->no_memops = e->no_memops || std::atoi(lb) != 0;
    if (const char *k = std::getenv("RAISR_CUDA_KERNEL")) e->use_pipe = std::strcmp(k, "tile") != 0;
    // 16-bit samples (not reachable from the FFmpeg filter: bits = 8..10, vf_raisr.c:83) run on the phase-sequential kernel: the
    // pipelined kernel reads its Gaussian weights from immutable constant tables that exist for 8 and 10 bit (raisr_gw_tables.h)
    if (cfg->bit_depth == 16) e->use_pipe = false;
    if (const char *c = std::getenv("RAISR_CUDA_CHAIN")) e->chain_passes = std::atoi(c) != 0 ? 1 : 0;
    if (const char *c = std::getenv("RAISR_CUDA_TEST_DROP_IN_FLAG")) e->test_drop_in_flag = std::atoi(c) != 0;
    if (const char *c = std::getenv("RAISR_CUDA_STAGE_PAGEABLE")) e->stage_pageable = std::atoi(c) != 0;
    if (const char *c = std::getenv("RAISR_CUDA_TMA")) e->use_tma = std::atoi(c) != 0;
    if (const char *c = std::getenv("RAISR_CUDA_COPY_THREADS")) e->copy_threads = std::max(0, std::min(16, std::atoi(c)));
    {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
        e->coop_launch = coop != 0;
    }
    if (e->cfg.numerics == RAISR_NUMERICS_FAST_HASH) {
        // opt-in fast numerics: separable structure tensor in front of the same eigen-analysis / quantisation
        if (cfg->bit_depth > 10 || !e->use_pipe) {
            std::cout << "[RAISR ERROR] the separable fast hash exists in the pipelined kernel (8 and 10 bit) only" << std::endl;
            return fail(RNLErrorBadParameter);
        }
        e->fast_hash = true;
        e->cfg.numerics = RAISR_NUMERICS_X86_IF_AVAILABLE;
    }
    if (e->cfg.numerics == RAISR_NUMERICS_FP16_FILTER) {
        // opt-in fast numerics: fp16 filter stage on top of the exact fp32 hash (buckets stay those of the fp32 path)
        if (cfg->bit_depth > 10) {
            std::cout << "[RAISR ERROR] the fp16 filter stage supports 8 and 10 bit only (16-bit samples overflow half precision)" << std::endl;
            return fail(RNLErrorBadParameter);
        }
        if (!e->use_pipe) {
            std::cout << "[RAISR ERROR] the fp16 filter stage exists in the pipelined kernel only" << std::endl;
            return fail(RNLErrorBadParameter);
        }
        e->filter_fp16 = true;
        e->cfg.numerics = RAISR_NUMERICS_X86_IF_AVAILABLE;
    }
    if (std::getenv("RAISR_CUDA_TIMING")) { e->timing = true; for (auto &ev : e->tev) cudaEventCreate(&ev); }
    for (unsigned i = 0; i < passes; ++i) {
        // device layout: [ptype][bucket][128], each row permuted so that the 8 lanes working on a pixel read 128
        // contiguous bytes per step: tap k = 16m + j  ->  position (m/2)*32 + (j/2)*4 + (m%2)*2 + (j%2)   (see dot8())
        const PassModel &pm = e->model.pass[i];
        if (pm.buckets > NBUCKET_MAX) {
            std::cout << "[RAISR ERROR] HashTable format is not compatible in number of hash keys!\n" << pm.buckets << std::endl;
            return fail(RNLErrorBadParameter);
        }
        std::vector<float> dev((size_t)pm.ptypes * pm.buckets * 128, 0.0f);
        for (int h = 0; h < pm.buckets; ++h)
            for (int t = 0; t < pm.ptypes; ++t)
                for (int k = 0; k < kTaps; ++k) {
                    const int m = k / 16, j = k % 16;
                    const int pos = (m / 2) * 32 + (j / 2) * 4 + (m % 2) * 2 + (j % 2);
                    dev[((size_t)t * pm.buckets + h) * 128 + pos] = pm.filters[((size_t)h * pm.ptypes + t) * kTapStride + k];
                }
        const size_t bytes = dev.size() * sizeof(float);
        if (cudaMalloc(&e->d_filters[i], bytes) != cudaSuccess ||
            cudaMemcpy(e->d_filters[i], dev.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess)
            return fail(RNLErrorInsufficientResources);
        if (e->filter_fp16) {                        // same layout, coefficients rounded to nearest-even binary16
            std::vector<__half> dev16(dev.size());
            for (size_t k = 0; k < dev.size(); ++k) dev16[k] = __float2half_rn(dev[k]);
            if (cudaMalloc(&e->d_filters16[i], dev16.size() * sizeof(__half)) != cudaSuccess ||
                cudaMemcpy(e->d_filters16[i], dev16.data(), dev16.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess)
                return fail(RNLErrorInsufficientResources);
        }
    }
    X86Tables xt{};
    const bool have_tables = x86_tables(&xt);
    if (e->cfg.numerics == RAISR_NUMERICS_X86_IF_AVAILABLE) e->cfg.numerics = have_tables ? RAISR_NUMERICS_X86 : RAISR_NUMERICS_IEEE;
    if (e->cfg.numerics == RAISR_NUMERICS_X86) {
        if (!have_tables) {
            std::cout << "[RAISR ERROR] x86 numerics tables are not built into this library" << std::endl;
            return fail(RNLErrorBadParameter);
        }
        // the 8-wide hash as compiled divides by 4 and by 2 through rcpps + one Newton step
        auto nr = [](float r, float den) { const float t = r * den; const float e2 = r * t; return (r + r) - e2; };
        e->quarter = nr(x86_rcpps(xt.rcpps, 4.0f), 4.0f);
        e->half = nr(x86_rcpps(xt.rcpps, 2.0f), 2.0f);
        // the kernels keep the two 14-bit tables packed as ((c0 >> 6) << 10) | c1 in shared memory (raisr_kernels.cuh: lut14)
        for (int t = 0; t < 2; ++t) {
            const uint32_t *tb = t ? xt.rcp14 : xt.rsqrt14;
            for (int i = 0; i < 128; ++i)
                if ((tb[2 * i] & 63u) != 0 || (tb[2 * i] >> 6) >= (1u << 22) || tb[2 * i + 1] >= 1024u) {
                    std::cout << "[RAISR ERROR] x86 numerics tables do not fit the packed form" << std::endl;
                    return fail(RNLErrorBadParameter);
                }
        }
        const void *src[4] = {xt.rsqrt14, xt.rcp14, xt.rsqrtps, xt.rcpps};
        const size_t bytes[4] = {kRsqrt14Words * sizeof(uint32_t), kRcp14Words * sizeof(uint32_t), kRsqrtpsEntries * sizeof(uint16_t),
                                 kRcppsEntries * sizeof(uint16_t)};
        for (int i = 0; i < 4; ++i)
            if (cudaMalloc(&e->d_lut[i], bytes[i]) != cudaSuccess ||
                cudaMemcpy(e->d_lut[i], src[i], bytes[i], cudaMemcpyHostToDevice) != cudaSuccess)
                return fail(RNLErrorInsufficientResources);
    }
    fill_weights(e, cfg->bit_depth);
    {
        // > 48 KB dynamic shared memory is a per-device function attribute: set for this engine's kernels on this engine's device
        int err = e->bps == 1 ? prepare_pass_tile<uint8_t>() : prepare_pass_tile<uint16_t>();
        if (!err) {
            const int nv = pipe_variant(e);
            if (e->bps == 1) err = nv == 1 ? prepare_frame_pipe<uint8_t, 1>() : nv == 2 ? prepare_frame_pipe<uint8_t, 2>() : nv == 4 ? prepare_frame_pipe<uint8_t, 4>() : prepare_frame_pipe<uint8_t, 0>();
            else err = nv == 1 ? prepare_frame_pipe<uint16_t, 1>() : nv == 2 ? prepare_frame_pipe<uint16_t, 2>() : nv == 4 ? prepare_frame_pipe<uint16_t, 4>() : prepare_frame_pipe<uint16_t, 0>();
        }
        if (err) { cuda_failed(err, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"); return fail(RNLErrorInsufficientResources); }
    }
    // flags and counters shared with the running kernel: all start at zero (the D2H stream waits for RUNNING totals of band_done)
    if (cudaMalloc(&e->d_chroma_ready, 2 * sizeof(unsigned)) != cudaSuccess || cudaMemset(e->d_chroma_ready, 0, 2 * sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc(&e->d_in_ready, sizeof(unsigned)) != cudaSuccess || cudaMemset(e->d_in_ready, 0, sizeof(unsigned)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->stream_h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&e->d_band_done, sizeof(unsigned) * raisr_cuda_engine::kMaxBands) != cudaSuccess ||
        cudaMemset(e->d_band_done, 0, sizeof(unsigned) * raisr_cuda_engine::kMaxBands) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->stream_d2h, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void **>(&e->h_err), sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess)
        return fail(RNLErrorInsufficientResources);
    *e->h_err = 0;
    if (cudaHostGetDevicePointer(reinterpret_cast<void **>(&e->d_err), e->h_err, 0) != cudaSuccess) return fail(RNLErrorInsufficientResources);
    {
        // cuStreamWaitValue32 through the runtime's driver entry point lookup (no link-time dependency on libcuda)
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            e->wait_value32 = reinterpret_cast<raisr_cuda_engine::WaitValue32Fn>(fn);
        else
            cudaGetLastError();
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            e->write_value32 = reinterpret_cast<raisr_cuda_engine::WaitValue32Fn>(fn);
        else
            cudaGetLastError();
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            e->encode_tiled = reinterpret_cast<raisr_cuda_engine::EncodeTiledFn>(fn);
        else
            cudaGetLastError();
    }
    for (auto &ev : e->ev_band)
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return fail(RNLErrorInsufficientResources);
    if (cudaEventCreateWithFlags(&e->ev_chroma_out, cudaEventDisableTiming) != cudaSuccess) return fail(RNLErrorInsufficientResources);
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->stream_uv, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_uv, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming) != cudaSuccess)
        return fail(RNLErrorInsufficientResources);
    *out = e;
    return RNLErrorNone;
}

int raisr_cuda_set_res(raisr_cuda_engine *e, unsigned in_w, unsigned in_h, unsigned out_w, unsigned out_h,
                       unsigned in_cw, unsigned in_ch, unsigned out_cw, unsigned out_ch)
{
    if (!e || !in_w || !in_h || !out_w || !out_h) return RNLErrorBadParameter;
    if (e->bind_device) CUDA_OK(cudaSetDevice(e->device));
    e->in_w = in_w; e->in_h = in_h; e->out_w = out_w; e->out_h = out_h;
    e->in_cw = in_cw; e->in_ch = in_ch; e->out_cw = out_cw; e->out_ch = out_ch;
    // the resize spec of the reference maps {inW, (int)(outH / ratio)} -> {outW, outH} (Raisr.cpp:1801-1803)
    int src_h = (int)((float)out_h / e->cfg.ratio);
    if (src_h > (int)in_h) src_h = in_h;
    if (src_h < 1) src_h = 1;
    e->yx.build(in_w, out_w);
    e->yy.build(src_h, out_h);
    e->up_src_h = src_h;
    if (e->yx.upload() || e->yy.upload()) return RNLErrorInsufficientResources;
    if (in_cw && in_ch && out_cw && out_ch) {
        e->cx.build(in_cw, out_cw);
        e->cy.build(in_ch, out_ch);
        if (e->cx.upload() || e->cy.upload()) return RNLErrorInsufficientResources;
    }
    if (e->d_in[0].alloc(in_w, in_h, e->bps) || e->d_out[0].alloc(out_w, out_h, e->bps)) return RNLErrorInsufficientResources;
    if (in_cw && in_ch && out_cw && out_ch)
        for (int i = 1; i < 3; ++i)
            if (e->d_in[i].alloc(in_cw, in_ch, e->bps) || e->d_out[i].alloc(out_cw, out_ch, e->bps))
                return RNLErrorInsufficientResources;
    if (e->cfg.passes == 2) {
        const bool mode2 = e->cfg.two_pass_mode == 2;       // intermediate is LR-sized in mode 2 (Raisr.cpp:1703-1723)
        if (e->d_mid.alloc(mode2 ? in_w : out_w, mode2 ? in_h : out_h, e->bps)) return RNLErrorInsufficientResources;
    }
    for (int i = 0; i < 3; ++i) {                                        // staging planes follow the geometry: reallocated on demand
        if (e->h_stage_in[i]) { cudaFreeHost(e->h_stage_in[i]); e->h_stage_in[i] = nullptr; }
        if (e->h_stage_out[i]) { cudaFreeHost(e->h_stage_out[i]); e->h_stage_out[i] = nullptr; }
    }
    cudaFree(e->d_rows_done); e->d_rows_done = nullptr; e->rows_done_cap = 0;
    if (e->cfg.passes == 2) {
        e->rows_done_cap = (int)std::max(in_h, out_h) / 2 + 2;          // tile rows of a pass (tiles are at least 2 rows high)
        CUDA_OK(cudaMalloc(&e->d_rows_done, sizeof(unsigned) * e->rows_done_cap));
    }
    for (int i = 0; i < 2; ++i) { cudaFree(e->d_hash[i]); e->d_hash[i] = nullptr; }
    if (e->cfg.keep_hash) {
        for (unsigned i = 0; i < e->cfg.passes; ++i) {
            const bool lr = e->cfg.passes == 2 && e->cfg.two_pass_mode == 2 && i == 0;
            e->hash_w[i] = lr ? in_w : out_w;
            e->hash_h[i] = lr ? in_h : out_h;
            CUDA_OK(cudaMalloc(&e->d_hash[i], sizeof(int) * (size_t)e->hash_w[i] * e->hash_h[i]));
        }
    }
    e->have_res = true;
    return RNLErrorNone;
}

int raisr_cuda_process_device_rows(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, void *out_y,
                                   size_t out_y_step, int blending, unsigned row0, unsigned row1, void *stream)
{
    if (!e || !e->have_res || !in_y || !out_y || row0 >= row1 || row1 > (unsigned)e->out_h) return RNLErrorBadParameter;
    if (check_blending(blending)) return RNLErrorBadParameter;
    e->blending = blending;
    e->sample_shift = 0;
    if (e->bind_device) CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (unsigned i = 0; i < e->cfg.passes; ++i)
        if (e->d_hash[i]) CUDA_OK(cudaMemsetAsync(e->d_hash[i], 0xff, sizeof(int) * (size_t)e->hash_w[i] * e->hash_h[i], s));
    return run_luma(e, in_y, in_y_step, out_y, out_y_step, (int)row0, (int)row1, s);
}

int raisr_cuda_process_device(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, const void *in_u,
                              size_t in_u_step, const void *in_v, size_t in_v_step, void *out_y, size_t out_y_step,
                              void *out_u, size_t out_u_step, void *out_v, size_t out_v_step, int blending, void *stream)
{
    if (!e || !e->have_res) return RNLErrorBadParameter;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (e->use_pipe && in_u && out_u && in_v && out_v && in_y && out_y) {
        // one launch per pass: the chroma planes ride along with the (first) luma launch
        if (check_blending(blending)) return RNLErrorBadParameter;
        e->blending = blending;
        e->sample_shift = 0;
        if (e->bind_device) CUDA_OK(cudaSetDevice(e->device));
        for (unsigned i = 0; i < e->cfg.passes; ++i)
            if (e->d_hash[i]) CUDA_OK(cudaMemsetAsync(e->d_hash[i], 0xff, sizeof(int) * (size_t)e->hash_w[i] * e->hash_h[i], s));
        ChromaJob cj{};
        cj.in[0] = in_u; cj.in[1] = in_v; cj.in_step[0] = in_u_step; cj.in_step[1] = in_v_step;
        cj.out[0] = out_u; cj.out[1] = out_v; cj.out_step[0] = out_u_step; cj.out_step[1] = out_v_step;
        return run_luma(e, in_y, in_y_step, out_y, out_y_step, 0, e->out_h, s, nullptr, nullptr, &cj);
    }
    int rc = raisr_cuda_process_device_rows(e, in_y, in_y_step, out_y, out_y_step, blending, 0, e->out_h, stream);
    if (rc) return rc;
    if (in_u && out_u) { rc = launch_resize(e, in_u, in_u_step, out_u, out_u_step, s); if (rc) return rc; }
    if (in_v && out_v) { rc = launch_resize(e, in_v, in_v_step, out_v, out_v_step, s); if (rc) return rc; }
    return RNLErrorNone;
}

}  // extern "C"

namespace {

// After a failed frame: nothing of it may leak into the next one.  The counters the copy streams wait on are running totals and
// the flags are sequence numbers, so a frame that stopped half way (an enqueue error, a timed-out flag wait) leaves them out of
// step with the host's shadow values: drain the device and start all of them from zero again.
void resync_after_failure(raisr_cuda_engine *e)
{
    cudaDeviceSynchronize();
    cudaGetLastError();
    if (e->pool) e->pool->wait_all();                                // copy jobs of the failed frame (their events have completed or failed by now)
    cudaMemset(e->d_band_done, 0, sizeof(unsigned) * raisr_cuda_engine::kMaxBands);
    cudaMemset(e->d_chroma_ready, 0, 2 * sizeof(unsigned));
    cudaMemset(e->d_in_ready, 0, sizeof(unsigned));
    cudaDeviceSynchronize();
    for (unsigned &t : e->band_target) t = 0;
    e->chroma_seq = e->chroma_done_target = e->frame_seq = 0;
    *e->h_err = 0;
}

int memop_failed(const char *what)
{
    std::cout << "[RAISR ERROR] " << what << " failed" << std::endl;
    return (int)RNLErrorUndefined;
}

// Phase-sequential kernel (RAISR_CUDA_KERNEL=tile, the cross-check implementation): everything in plain stream order, chroma
// on its own stream with one resize launch per plane.
int process_host_tile(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, const void *in_u, size_t in_u_step, const void *in_v,
                      size_t in_v_step, void *out_y, size_t out_y_step, void *out_u, size_t out_u_step, void *out_v, size_t out_v_step,
                      bool chroma)
{
    const size_t bps = e->bps;
    if (chroma) {
        const void *src[2] = {in_u, in_v};
        const size_t sstep[2] = {in_u_step, in_v_step};
        void *dst[2] = {out_u, out_v};
        const size_t dstep[2] = {out_u_step, out_v_step};
        for (int i = 0; i < 2; ++i) {
            CUDA_OK(cudaMemcpy2DAsync(e->d_in[i + 1].ptr, e->d_in[i + 1].pitch, src[i], sstep[i], e->in_cw * bps, e->in_ch,
                                      cudaMemcpyHostToDevice, e->stream_uv));
            int rc = launch_resize(e, e->d_in[i + 1].ptr, e->d_in[i + 1].pitch, e->d_out[i + 1].ptr, e->d_out[i + 1].pitch, e->stream_uv);
            if (rc) return rc;
        }
        for (int i = 0; i < 2; ++i)
            CUDA_OK(cudaMemcpy2DAsync(dst[i], dstep[i], e->d_out[i + 1].ptr, e->d_out[i + 1].pitch, e->out_cw * bps, e->out_ch,
                                      cudaMemcpyDeviceToHost, e->stream_uv));
    }
    CUDA_OK(cudaMemcpy2DAsync(e->d_in[0].ptr, e->d_in[0].pitch, in_y, in_y_step, e->in_w * bps, e->in_h, cudaMemcpyHostToDevice, e->stream));
    for (unsigned i = 0; i < e->cfg.passes; ++i)
        if (e->d_hash[i]) CUDA_OK(cudaMemsetAsync(e->d_hash[i], 0xff, sizeof(int) * (size_t)e->hash_w[i] * e->hash_h[i], e->stream));
    int rc = run_luma(e, e->d_in[0].ptr, e->d_in[0].pitch, e->d_out[0].ptr, e->d_out[0].pitch, 0, e->out_h, e->stream);
    if (rc) return rc;
    CUDA_OK(cudaMemcpy2DAsync(out_y, out_y_step, e->d_out[0].ptr, e->d_out[0].pitch, e->out_w * bps, e->out_h, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    if (chroma) CUDA_OK(cudaStreamSynchronize(e->stream_uv));
    return RNLErrorNone;
}

// Pipelined kernel: ONE launch per pass carries the whole frame (the chroma planes ride along in the filter warps); the copies run
// on three side streams under the kernel, ordered against it by flags (stream memory operations):
//   luma in   : the first rows ahead of the launch, the rest on stream_h2d, flag in_ready  -> kernel (both reader groups wait)
//   chroma in : behind the luma copies on stream_uv,              flag chroma_ready        -> kernel (filter warps wait)
//   chroma out: kernel counts CTAs done with their share (chroma_done) -> stream_uv waits, copies the planes out
//   luma out  : kernel counts finished tiles per row band (band_done)  -> stream_d2h waits per band, copies the band; the rows of
//               the last round of tiles are written in place when the caller's plane is page-locked
// Every copy the kernel waits for is enqueued BEFORE the launch.  Without stream memory operations (or RAISR_CUDA_NO_MEMOPS=1,
// needed under tools that serialise the GPU) the same copies run in plain stream order around the kernel.
int process_host_pipe(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, const void *in_u, size_t in_u_step, const void *in_v,
                      size_t in_v_step, void *out_y, size_t out_y_step, void *out_u, size_t out_u_step, void *out_v, size_t out_v_step,
                      bool chroma)
{
    const size_t bps = e->bps;
    const bool memops = e->wait_value32 && e->write_value32 && !e->no_memops;
    const auto t_entry = std::chrono::steady_clock::now();
    auto mark = [&](int i) { if (e->timing && e->t_n >= 8) e->t_host[i] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_entry).count(); };

    // ---- pageable caller planes: go through page-locked staging planes, bytes moved by the copy threads ------------------
    const void *probe = nullptr;
    const bool stage_in = e->stage_pageable && !mapped_host_pointer(in_y, &probe);
    const bool stage_out = e->stage_pageable && !mapped_host_pointer(out_y, &probe);
    void *user_out[3] = {out_y, out_u, out_v};
    const size_t user_ostep[3] = {out_y_step, out_u_step, out_v_step};
    const size_t orow[3] = {e->out_w * bps, e->out_cw * bps, e->out_cw * bps};
    const int orows[3] = {e->out_h, e->out_ch, e->out_ch};
    bool late_inputs = false;                                           // staged input rows that are still being copied at launch time
    if (stage_in || stage_out) {
        if (!e->pool) { e->pool = new CopyPool; e->pool->start(e->copy_threads, e->device, e->bind_device); }
        const size_t irow[3] = {e->in_w * bps, e->in_cw * bps, e->in_cw * bps};
        const int irows[3] = {e->in_h, e->in_ch, e->in_ch};
        for (int i = 0; i < (chroma ? 3 : 1); ++i) {
            if (stage_in && !e->h_stage_in[i]) CUDA_OK(cudaHostAlloc(&e->h_stage_in[i], irow[i] * irows[i], cudaHostAllocDefault));
            if (stage_out && !e->h_stage_out[i]) CUDA_OK(cudaHostAlloc(&e->h_stage_out[i], orow[i] * orows[i], cudaHostAllocDefault));
        }
        if (stage_in) {
            // Without stream memory operations every input byte is staged before anything is enqueued.  With them only the luma
            // rows the first tiles read are staged up front; the copy threads stage the rest (~3 MB) while this thread enqueues
            // the first H2D copy and launches the kernel, which waits in-kernel (bounded) for the flags behind the late copies.
            const void *src[3] = {in_y, in_u, in_v};
            const size_t sstep[3] = {in_y_step, in_u_step, in_v_step};
            late_inputs = memops && e->split_h2d && e->in_h >= 256 && e->in_h < 65536;
            const int early_rows = late_inputs ? std::max(64, (e->in_h / 8 + 15) & ~15) : 0;
            // luma rows first (group 1: their H2D copy is what the running kernel waits for), then the chroma planes (group 2),
            // every plane cut so that all copy threads have work
            const int parts = std::max(1, e->copy_threads);
            auto stage_rows = [&](int i, int r0, int r1, int group) {
                if (r1 <= r0) return;
                void *dst = static_cast<char *>(e->h_stage_in[i]) + (size_t)r0 * irow[i];
                const void *sp = static_cast<const char *>(src[i]) + (size_t)r0 * sstep[i];
                const size_t ss = sstep[i], rb = irow[i];
                if (group < 0) copy_rows(dst, rb, sp, ss, rb, r1 - r0);
                else e->pool->submit(nullptr, [dst, sp, ss, rb, r0, r1] { copy_rows(dst, rb, sp, ss, rb, r1 - r0); }, group);
            };
            // the early rows are on the critical path (nothing runs until they are on the device): all copy threads and this one
            for (int k = 1; k <= parts && early_rows; ++k)
                stage_rows(0, (int)((long long)early_rows * k / (parts + 1)), (int)((long long)early_rows * (k + 1) / (parts + 1)), 0);
            for (int i = 0; i < (chroma ? 3 : 1); ++i) {
                const int first = i == 0 ? early_rows : 0, n = i == 0 ? parts : std::max(1, parts / 2);
                for (int k = 0; k < n; ++k)
                    stage_rows(i, first + (int)((long long)(irows[i] - first) * k / n), first + (int)((long long)(irows[i] - first) * (k + 1) / n), i == 0 ? 1 : 2);
            }
            if (early_rows) {
                stage_rows(0, 0, early_rows / (parts + 1), -1);
                e->pool->wait_group(0);
            }
            if (!late_inputs) e->pool->wait_all();
            in_y = e->h_stage_in[0]; in_y_step = irow[0];
            if (chroma) { in_u = e->h_stage_in[1]; in_v = e->h_stage_in[2]; in_u_step = in_v_step = irow[1]; }
        }
        if (stage_out) {
            out_y = e->h_stage_out[0]; out_y_step = orow[0];
            if (chroma) { out_u = e->h_stage_out[1]; out_v = e->h_stage_out[2]; out_u_step = out_v_step = orow[1]; }
        }
    }
    // rows [r0, r1) of staged output plane i -> the caller's plane, once `ev` has completed (by a copy thread)
    auto deliver = [&](int i, int r0, int r1, cudaEvent_t ev) {
        char *dst = static_cast<char *>(user_out[i]) + (size_t)r0 * user_ostep[i];
        const char *src = static_cast<const char *>(e->h_stage_out[i]) + (size_t)r0 * orow[i];
        const size_t ds = user_ostep[i], rb = orow[i];
        e->pool->submit(ev, [dst, src, ds, rb, r0, r1] { copy_rows(dst, ds, src, rb, rb, r1 - r0); });
    };

    const void *csrc[2] = {in_u, in_v};
    const size_t csstep[2] = {in_u_step, in_v_step};
    void *cdst[2] = {out_u, out_v};
    const size_t cdstep[2] = {out_u_step, out_v_step};

    // ---- luma input: rows [0, split_row) ahead of the launch, the rest under the kernel ----------------------------------
    int split_row = 0;
    const unsigned *in_ready = nullptr;
    if (memops && e->split_h2d && e->in_h >= 256 && e->in_h < 65536) {
        split_row = std::max(64, (e->in_h / 8 + 15) & ~15);
        ++e->frame_seq;
        in_ready = e->d_in_ready;
    }
    const int rows0 = split_row ? split_row : e->in_h;
    mark(0);                                                                // early rows staged
    if (e->timing) cudaEventRecord(e->tev[0], e->stream);
    CUDA_OK(cudaMemcpy2DAsync(e->d_in[0].ptr, e->d_in[0].pitch, in_y, in_y_step, e->in_w * bps, rows0, cudaMemcpyHostToDevice, e->stream));
    if (e->timing) cudaEventRecord(e->tev[1], e->stream);
    // luma rows >= split_row and the chroma planes: on their own streams, each followed by the flag the kernel waits for.  Enqueued
    // before the launch -- unless their staging copies are still running (late_inputs): then right after it.
    ChromaJob cj{};
    if (chroma) {
        for (int i = 0; i < 2; ++i) {
            cj.in[i] = e->d_in[i + 1].ptr; cj.in_step[i] = e->d_in[i + 1].pitch;
            cj.out[i] = e->d_out[i + 1].ptr; cj.out_step[i] = e->d_out[i + 1].pitch;
        }
        if (memops) {
            cj.ready = e->d_chroma_ready; cj.seq = ++e->chroma_seq;          // H2D on the chroma stream, flagged to the running kernel
            cj.done = e->d_chroma_ready + 1;                                 // D2H by the copy engine as soon as every CTA has written its share
        }
    }
    if (split_row) CUDA_OK(cudaEventRecord(e->ev_uv, e->stream));           // part 2 behind part 1 (same copy engine anyway)
    // (test hook RAISR_CUDA_TEST_DROP_IN_FLAG=1: the first frame's watermarks are never written -- what a failed copy looks like to
    //  the kernel; exercises the bounded spin, the error word and the fall-back to plain stream order)
    const bool drop_flags = e->test_drop_in_flag;
    e->test_drop_in_flag = false;
    // The rest of the plane in ONE copy followed by the watermark write ((frame << 16) | rows that have landed).  (The kernel's
    // protocol allows several flagged copies; measured, they only cost: 4 copies 1670 vs 1765 frames/s at N = 1, 7689 vs 7778 on
    // 8 GPUs whose host link is slower than the kernel, 1545 vs 1580 with pageable planes -- DESIGN.md section 5.)
    auto enqueue_rest_of_luma = [&]() -> int {
        if (!split_row) return 0;
        CUDA_OK(cudaStreamWaitEvent(e->stream_h2d, e->ev_uv, 0));
        CUDA_OK(cudaMemcpy2DAsync(static_cast<char *>(e->d_in[0].ptr) + (size_t)split_row * e->d_in[0].pitch, e->d_in[0].pitch,
                                  static_cast<const char *>(in_y) + (size_t)split_row * in_y_step, in_y_step, e->in_w * bps, e->in_h - split_row,
                                  cudaMemcpyHostToDevice, e->stream_h2d));
        if (!drop_flags && e->write_value32(e->stream_h2d, (unsigned long long)(uintptr_t)e->d_in_ready, (e->frame_seq << 16) | (unsigned)e->in_h, 0) != 0)
            return memop_failed("cuStreamWriteValue32");
        return 0;
    };
    auto record_luma_in = [&]() -> int {                                    // the chroma copies queue up behind the luma copies
        if (chroma) CUDA_OK(cudaEventRecord(e->ev_in, split_row ? e->stream_h2d : e->stream));
        return 0;
    };
    auto enqueue_chroma_in = [&]() -> int {
        if (chroma) {
            if (memops) {
                CUDA_OK(cudaStreamWaitEvent(e->stream_uv, e->ev_in, 0));
                for (int i = 0; i < 2; ++i)
                    CUDA_OK(cudaMemcpy2DAsync(e->d_in[i + 1].ptr, e->d_in[i + 1].pitch, csrc[i], csstep[i], e->in_cw * bps, e->in_ch,
                                              cudaMemcpyHostToDevice, e->stream_uv));
                if (e->write_value32(e->stream_uv, (unsigned long long)(uintptr_t)e->d_chroma_ready, cj.seq, 0) != 0) return memop_failed("cuStreamWriteValue32");
            } else {
                for (int i = 0; i < 2; ++i)
                    CUDA_OK(cudaMemcpy2DAsync(e->d_in[i + 1].ptr, e->d_in[i + 1].pitch, csrc[i], csstep[i], e->in_cw * bps, e->in_ch,
                                              cudaMemcpyHostToDevice, e->stream));
            }
        }
        return 0;
    };
    // Every copy the kernel waits for is enqueued ahead of the launch -- unless the staging copies of a pageable plane are still
    // running (late_inputs): then right behind it.
    if (!late_inputs) {
        int rc_in = enqueue_rest_of_luma();
        if (!rc_in) rc_in = record_luma_in();
        if (!rc_in) rc_in = enqueue_chroma_in();
        if (rc_in) return rc_in;
    }
    for (unsigned i = 0; i < e->cfg.passes; ++i)
        if (e->d_hash[i]) CUDA_OK(cudaMemsetAsync(e->d_hash[i], 0xff, sizeof(int) * (size_t)e->hash_w[i] * e->hash_h[i], e->stream));

    // ---- launch ------------------------------------------------------------------------------------------------------------
    const bool band_d2h = memops && e->band_pipeline;
    const void *tail_dev = nullptr;                                         // rows of the last round of tiles: in place when the plane is page-locked
    const bool tail_direct = band_d2h && e->tail_in_place && mapped_host_pointer(out_y, &tail_dev);
    int rc = run_luma(e, e->d_in[0].ptr, e->d_in[0].pitch, e->d_out[0].ptr, e->d_out[0].pitch, 0, e->out_h, e->stream,
                      band_d2h ? e->d_band_done : nullptr, in_ready, chroma ? &cj : nullptr,
                      tail_direct ? const_cast<void *>(tail_dev) : nullptr, out_y_step, split_row);
    if (rc) return rc;
    if (e->timing) cudaEventRecord(e->tev[2], e->stream);
    mark(1);                                                                // kernel launched
    if (late_inputs) {                                                      // the kernel is running on the early rows: now the rest
        e->pool->wait_group(1);                                             // luma rows staged: their copy and watermark first ...
        int rc_in = enqueue_rest_of_luma();
        if (!rc_in) rc_in = record_luma_in();
        if (rc_in) return rc_in;
        e->pool->wait_group(2);                                             // ... the chroma planes (needed from the third tile on) behind them
        mark(2);                                                            // late input rows staged
        rc_in = enqueue_chroma_in();
        if (rc_in) return rc_in;
    }

    // ---- chroma output -----------------------------------------------------------------------------------------------------
    if (chroma) {
        cudaStream_t cs = e->stream;
        if (memops) {
            cs = e->stream_uv;
            e->chroma_done_target += (unsigned)e->last_chroma_ctas;          // running total: no reset, no race with the previous frame
            if (e->wait_value32(cs, (unsigned long long)(uintptr_t)(e->d_chroma_ready + 1), e->chroma_done_target, 0 /* GEQ */) != 0)
                return memop_failed("cuStreamWaitValue32");
        }
        for (int i = 0; i < 2; ++i)
            CUDA_OK(cudaMemcpy2DAsync(cdst[i], cdstep[i], e->d_out[i + 1].ptr, e->d_out[i + 1].pitch, e->out_cw * bps, e->out_ch,
                                      cudaMemcpyDeviceToHost, cs));
        if (stage_out) CUDA_OK(cudaEventRecord(e->ev_chroma_out, cs));
    }
    // (queued BEHIND the luma bands: the pool takes the oldest job whose event has completed and sleeps on the oldest one when none
    //  has -- which must be the first band, not the chroma planes that the kernel finishes around its middle)
    auto deliver_chroma = [&]() {
        if (!(chroma && stage_out)) return;
        const int parts = std::max(1, e->copy_threads / 2);                     // 2 MB per plane at 4K: not one thread's job
        for (int i = 1; i <= 2; ++i)
            for (int k = 0; k < parts; ++k) {
                const int a = (int)((long long)e->out_ch * k / parts), b = (int)((long long)e->out_ch * (k + 1) / parts);
                if (b > a) deliver(i, a, b, e->ev_chroma_out);
            }
    };

    // ---- luma output -------------------------------------------------------------------------------------------------------
    if (band_d2h) {
        // The final pass counts finished tiles per row band (running totals, never reset); the D2H stream waits on each counter
        // and copies that band while the kernel is still working on the rows below (copies overlap compute inside ONE frame).
        for (int b = 0; b < e->n_bands; ++b) {
            e->band_target[b] += e->band_tiles[b];
            if (e->wait_value32(e->stream_d2h, (unsigned long long)(uintptr_t)(e->d_band_done + b), e->band_target[b], 0 /* GEQ */) != 0)
                return memop_failed("cuStreamWaitValue32");
            const int r0 = e->band_row0[b], r1 = e->band_row1[b];
            CUDA_OK(cudaMemcpy2DAsync(static_cast<char *>(out_y) + (size_t)r0 * out_y_step, out_y_step,
                                      static_cast<char *>(e->d_out[0].ptr) + (size_t)r0 * e->d_out[0].pitch, e->d_out[0].pitch,
                                      e->out_w * bps, r1 - r0, cudaMemcpyDeviceToHost, e->stream_d2h));
            if (stage_out) {
                CUDA_OK(cudaEventRecord(e->ev_band[b], e->stream_d2h));
                if (b + 2 >= e->n_bands && r1 - r0 >= 2) {                      // the last bands are on the critical path: two threads each
                    deliver(0, r0, (r0 + r1) / 2, e->ev_band[b]);
                    deliver(0, (r0 + r1) / 2, r1, e->ev_band[b]);
                } else {
                    deliver(0, r0, r1, e->ev_band[b]);
                }
            }
        }
        if (!tail_direct && e->n_bands > 0 && e->band_row1[e->n_bands - 1] < e->out_h) {
            // (not reached today: without an in-place tail every tile row belongs to a band)
            const int r0 = e->band_row1[e->n_bands - 1];
            CUDA_OK(cudaMemcpy2DAsync(static_cast<char *>(out_y) + (size_t)r0 * out_y_step, out_y_step,
                                      static_cast<char *>(e->d_out[0].ptr) + (size_t)r0 * e->d_out[0].pitch, e->d_out[0].pitch,
                                      e->out_w * bps, e->out_h - r0, cudaMemcpyDeviceToHost, e->stream));
        }
        deliver_chroma();
        mark(3);                                                            // everything enqueued
        CUDA_OK(cudaStreamSynchronize(e->stream_d2h));
        mark(4);                                                            // last band copied out
    } else {
        deliver_chroma();
        CUDA_OK(cudaMemcpy2DAsync(out_y, out_y_step, e->d_out[0].ptr, e->d_out[0].pitch, e->out_w * bps, e->out_h, cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    mark(5);                                                                // kernel done
    if (stage_out) {
        // what the bands did not cover: the rows of the last round of tiles (written in place into the staging plane), or the whole
        // plane when there is no band pipeline
        const int r0 = band_d2h ? (e->n_bands > 0 ? e->band_row1[e->n_bands - 1] : 0) : 0;
        const int parts = e->copy_threads + 1;                              // the copy threads and this one
        for (int k = 1; k < parts && r0 < e->out_h; ++k) {
            const int a = r0 + (int)((long long)(e->out_h - r0) * k / parts), b = r0 + (int)((long long)(e->out_h - r0) * (k + 1) / parts);
            if (b > a) deliver(0, a, b, nullptr);
        }
        if (r0 < e->out_h) {
            const int b = r0 + (int)((long long)(e->out_h - r0) / parts);
            if (b > r0) copy_rows(static_cast<char *>(user_out[0]) + (size_t)r0 * user_ostep[0], user_ostep[0],
                                  static_cast<const char *>(e->h_stage_out[0]) + (size_t)r0 * orow[0], orow[0], orow[0], b - r0);
        }
        e->pool->wait_all();
    }
    mark(6);                                                                // every staged row delivered
    if (split_row) CUDA_OK(cudaStreamSynchronize(e->stream_h2d));
    if (chroma && memops) CUDA_OK(cudaStreamSynchronize(e->stream_uv));
    if (e->timing) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e->tev[0], e->tev[1]); cudaEventElapsedTime(&b, e->tev[1], e->tev[2]);
        e->t_h2d += a; e->t_kern += b; e->t_n++;
    }
    mark(7);
    return RNLErrorNone;
}

}  // namespace

extern "C" {

int raisr_cuda_process_device_semiplanar(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, const void *in_uv, size_t in_uv_step,
                                         void *out_y, size_t out_y_step, void *out_uv, size_t out_uv_step, int sample_shift, int blending,
                                         void *stream)
{
    if (!e || !e->have_res || !in_y || !out_y || !in_uv || !out_uv) return RNLErrorBadParameter;
    if (sample_shift < 0 || sample_shift > 8 || (e->bps == 1 && sample_shift != 0)) return RNLErrorBadParameter;
    if (!e->use_pipe) {
        std::cout << "[RAISR ERROR] semi-planar frames need the pipelined kernel" << std::endl;
        return RNLErrorBadParameter;
    }
    if (check_blending(blending)) return RNLErrorBadParameter;
    e->blending = blending;
    e->sample_shift = sample_shift;
    if (e->bind_device) CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (unsigned i = 0; i < e->cfg.passes; ++i)
        if (e->d_hash[i]) CUDA_OK(cudaMemsetAsync(e->d_hash[i], 0xff, sizeof(int) * (size_t)e->hash_w[i] * e->hash_h[i], s));
    ChromaJob cj{};
    cj.comps = 2;
    cj.in[0] = in_uv; cj.in_step[0] = in_uv_step;
    cj.out[0] = out_uv; cj.out_step[0] = out_uv_step;
    const int rc = run_luma(e, in_y, in_y_step, out_y, out_y_step, 0, e->out_h, s, nullptr, nullptr, &cj);
    e->sample_shift = 0;
    return rc;
}

int raisr_cuda_process_host(raisr_cuda_engine *e, const void *in_y, size_t in_y_step, const void *in_u,
                            size_t in_u_step, const void *in_v, size_t in_v_step, void *out_y, size_t out_y_step,
                            void *out_u, size_t out_u_step, void *out_v, size_t out_v_step, int blending)
{
    if (!e || !e->have_res || !in_y || !out_y) return RNLErrorBadParameter;
    if (check_blending(blending)) return RNLErrorBadParameter;
    e->blending = blending;
    e->sample_shift = 0;
    if (e->bind_device) CUDA_OK(cudaSetDevice(e->device));
    const bool chroma = in_u && in_v && out_u && out_v && e->d_in[1].ptr;
    int rc = e->use_pipe ? process_host_pipe(e, in_y, in_y_step, in_u, in_u_step, in_v, in_v_step, out_y, out_y_step, out_u, out_u_step, out_v, out_v_step, chroma)
                         : process_host_tile(e, in_y, in_y_step, in_u, in_u_step, in_v, in_v_step, out_y, out_y_step, out_u, out_u_step, out_v, out_v_step, chroma);
    if (rc == RNLErrorNone && *reinterpret_cast<volatile unsigned *>(e->h_err) != 0) {
        // An in-kernel flag wait ran into its bound: the copy behind the flag never made progress while the kernel was running.
        // That is what tools that serialise the GPU do (ncu, compute-sanitizer, CUDA_LAUNCH_BLOCKING): re-run the frame in plain
        // stream order and stay there for the rest of this engine's life.  A second failure is an error.
        resync_after_failure(e);
        if (e->use_pipe && !e->no_memops) {
            std::cout << "[RAISR WARNING] a copy the kernel was waiting for made no progress under the running kernel (profiler / "
                         "serialised GPU?): switching this engine to plain stream order (RAISR_CUDA_NO_MEMOPS=1)" << std::endl;
            e->no_memops = true;
            rc = process_host_pipe(e, in_y, in_y_step, in_u, in_u_step, in_v, in_v_step, out_y, out_y_step, out_u, out_u_step, out_v, out_v_step, chroma);
            if (rc == RNLErrorNone && *reinterpret_cast<volatile unsigned *>(e->h_err) != 0) rc = RNLErrorUndefined;
        } else {
            rc = RNLErrorUndefined;
        }
        if (rc != RNLErrorNone) std::cout << "[RAISR ERROR] a copy the kernel was waiting for never arrived (flag wait timed out)" << std::endl;
    }
    if (rc != RNLErrorNone) resync_after_failure(e);
    return rc;
}

int raisr_cuda_ipc_export(void *device_ptr, unsigned char handle[64], size_t *offset)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size of the C ABI");
    if (!device_ptr || !handle || !offset) return RNLErrorBadParameter;
    // base of the allocation device_ptr lies in (cuMemGetAddressRange through the runtime's driver entry point lookup)
    typedef int (*GetRangeFn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    unsigned long long base = 0;
    size_t size = 0;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess ||
        reinterpret_cast<GetRangeFn>(fn)(&base, &size, (unsigned long long)(uintptr_t)device_ptr) != 0) {
        cudaGetLastError();
        return RNLErrorUndefined;
    }
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>((uintptr_t)base)));
    memcpy(handle, &h, sizeof(h));
    *offset = (size_t)((uintptr_t)device_ptr - (uintptr_t)base);
    return RNLErrorNone;
}

int raisr_cuda_ipc_open(const unsigned char handle[64], size_t offset, void **device_ptr)
{
    if (!handle || !device_ptr) return RNLErrorBadParameter;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *base = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *device_ptr = static_cast<char *>(base) + offset;
    return RNLErrorNone;
}

int raisr_cuda_ipc_close(void *device_ptr)
{
    // (the pointer returned by raisr_cuda_ipc_open minus its offset: callers keep the pair; closing an interior pointer is an error)
    if (!device_ptr) return RNLErrorBadParameter;
    CUDA_OK(cudaIpcCloseMemHandle(device_ptr));
    return RNLErrorNone;
}

int raisr_cuda_read_hash(raisr_cuda_engine *e, int pass, int32_t *host_out, size_t count)
{
    if (!e || pass < 0 || pass > 1 || !e->d_hash[pass] || !host_out) return RNLErrorBadParameter;
    if (count != (size_t)e->hash_w[pass] * e->hash_h[pass]) return RNLErrorBadParameter;
    if (e->bind_device) CUDA_OK(cudaSetDevice(e->device));
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(host_out, e->d_hash[pass], count * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return RNLErrorNone;
}

unsigned long long raisr_cuda_launch_count(const raisr_cuda_engine *e) { return e ? e->launches : 0; }

int raisr_cuda_numerics(const raisr_cuda_engine *e)
{
    if (!e) return -1;
    return e->filter_fp16 ? (int)RAISR_NUMERICS_FP16_FILTER : e->fast_hash ? (int)RAISR_NUMERICS_FAST_HASH : e->cfg.numerics;
}

void raisr_cuda_destroy(raisr_cuda_engine *e)
{
    if (!e) return;
    if (e->bind_device) cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    const double tn = e->t_n > 8 ? double(e->t_n - 8) : 1.0;      // host marks skip the first 8 frames (allocation of the staging planes)
    if (e->timing && e->t_n) std::cout << "[RAISR TIMING] frames " << e->t_n << " luma H2D " << 1e3 * e->t_h2d / e->t_n << " us, memset+kernel " << 1e3 * e->t_kern / e->t_n << " us; host marks (us since entry): early rows staged " << e->t_host[0] / tn
                                        << ", launched " << e->t_host[1] / tn << ", late rows staged " << e->t_host[2] / tn << ", all enqueued " << e->t_host[3] / tn
                                        << ", bands out " << e->t_host[4] / tn << ", kernel done " << e->t_host[5] / tn << ", delivered " << e->t_host[6] / tn
                                        << ", return " << e->t_host[7] / tn << std::endl;
    for (int i = 0; i < 2; ++i) { cudaFree(e->d_filters[i]); cudaFree(e->d_filters16[i]); cudaFree(e->d_hash[i]); }
    cudaFree(e->d_rows_done);
    for (int i = 0; i < 4; ++i) cudaFree(e->d_lut[i]);
    for (int i = 0; i < 3; ++i) { e->d_in[i].release(); e->d_out[i].release(); }
    e->d_mid.release();
    e->yx.release(); e->yy.release(); e->cx.release(); e->cy.release();
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->stream_d2h) cudaStreamDestroy(e->stream_d2h);
    if (e->stream_h2d) cudaStreamDestroy(e->stream_h2d);
    cudaFree(e->d_in_ready);
    cudaFree(e->d_chroma_ready);
    cudaFree(e->d_band_done);
    if (e->h_err) cudaFreeHost(e->h_err);
    delete e->pool;
    for (int i = 0; i < 3; ++i) { if (e->h_stage_in[i]) cudaFreeHost(e->h_stage_in[i]); if (e->h_stage_out[i]) cudaFreeHost(e->h_stage_out[i]); }
    for (auto &ev : e->ev_band) if (ev) cudaEventDestroy(ev);
    if (e->ev_chroma_out) cudaEventDestroy(e->ev_chroma_out);
    if (e->stream_uv) cudaStreamDestroy(e->stream_uv);
    if (e->ev_uv) cudaEventDestroy(e->ev_uv);
    if (e->ev_in) cudaEventDestroy(e->ev_in);
    delete e;
}

}  // extern "C"
