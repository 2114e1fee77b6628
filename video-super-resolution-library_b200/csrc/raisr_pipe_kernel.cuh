// raisr_pipe_kernel.cuh -- persistent, warp-specialised form of the RAISR pass kernel (the default).
//
// Same arithmetic as raisr_pass_kernel (raisr_kernels.cuh: stages A-E, bit-identical results); what changes is the
// schedule.  The profile of the phase-sequential kernel shows two kinds of stages: B/C (structure tensor + bucket) are
// bound by FP32/ALU issue with the shared-memory pipe idle, D (121-tap filter out of the shared filter slice) is bound
// by the shared-memory pipe with the FMA pipe idle.  Here one CTA per SM loops over tiles with three warp roles:
//   chain warps  (8) : tile i+1 -- stream the upscaled rows through a 16-row ring, stage B column chains (FMUL2/FFMA2)
//   bucket warps (8) : tile i+1 -- stage C: lane tree, eigen-analysis, hash              -> bucket tile[(i+1)&1]
//   filter warps (12): tile i   -- S tile, per-type filter slices (cp.async.bulk), 8-lane filter, blend, store
// so the FMA-bound, the latency-bound and the LSU-bound work overlap on the same SM.  The chain warps run one 2-row chunk
// ahead of the bucket warps through a double-buffered chain buffer (one named barrier per chunk); producers and filter
// warps hand over a pair of bucket tiles guarded by hardware named barriers (bar.arrive / bar.sync, full and empty per
// tile buffer: the waiting side blocks without consuming issue slots).
//
// Register file: the CTA is launched with 72 registers per thread; the chain and the bucket warpgroups drop to 56, the filter
// warpgroups rise to 88 (setmaxnreg; 48 / 56 / 96 while the filter warps were the bottleneck).  The upscaled rows of chunk k+2 are fetched from global memory before the barrier
// of chunk k and stored into free ring slots after stage B (latency hidden).
//
// Round 2: stage D walks down pixel columns with the patch values in registers (sliding window, rotating chain ownership: see
// pipe_filter_pass); two passes can be chained in one cooperative launch.  The kernel is bound by instruction issue.
//
// STATUS: bit-identical to raisr_pass_kernel (same GPU parity suite, tools/kbench.py); 0.505 ms vs 0.85 ms per 4K frame.
#pragma once
#include "raisr_kernels.cuh"
#include "raisr_gw_tables.h"

namespace raisr {

constexpr int NTP = 896;                     // threads per CTA of the pipelined kernel: 12 filter + 8 chain + 8 bucket warps
constexpr int NPW = 16;                      // producer warps: 8 chain warps (stage B) + 8 bucket warps (stage C)
constexpr int NCW = NTP / 32 - NPW;          // filter (consumer) warps (3 warpgroups)
constexpr int NPT = NPW * 32, NCT = NCW * 32;
constexpr int NBT = NPT / 2;                 // threads of each producer sub-role
#ifndef RAISR_CHAIN_REGS
#define RAISR_CHAIN_REGS 56
#endif
#ifndef RAISR_BUCKET_REGS
#define RAISR_BUCKET_REGS 56
#endif
#ifndef RAISR_CONS_REGS
#define RAISR_CONS_REGS 88
#endif
constexpr int CHAIN_REGS = RAISR_CHAIN_REGS, BUCKET_REGS = RAISR_BUCKET_REGS, CONS_REGS = RAISR_CONS_REGS;   // setmaxnreg targets: 256*56 + 256*56 + 384*88 <= 896*72 registers of the CTA
                                                                   // (measured, round 1: 48/48/104 0.615 ms, 56/56/88 0.606 ms, 64/64/80 0.615 ms, 48/64/88 0.600 ms; round 2: DESIGN.md section 4)
static_assert(NBT * (CHAIN_REGS + BUCKET_REGS) + NCT * CONS_REGS <= NTP * 72, "register file of the CTA");
constexpr int RBP = 2;                       // filtered rows per producer chunk (RBP * QW == NBT positions)
constexpr int RING = 16;                     // rows of the producer's S ring (>= RBP + 12, power of two)
static_assert(RBP * QW == NBT && RING == 2 * RBP + 12 && (RING & (RING - 1)) == 0 && RBP == 2 && SW % 2 == 0, "producer geometry");

constexpr int PTH_MAX = 46;                  // output tile height of the pipelined kernel (smaller than TH_MAX: the chain buffer is double-buffered)
constexpr int PSH = PTH_MAX + 14, PHH = PTH_MAX + 2;
constexpr int QCHUNK = RBP * 18 * QW;        // floats of one chunk's column chains
constexpr size_t POFF_S = 0;
constexpr size_t POFF_HR = POFF_S + sizeof(float) * PSH * SP;
constexpr size_t POFF_F = (POFF_HR + sizeof(float) * PHH * HP + 127) & ~(size_t)127;
constexpr size_t POFF_HASH = POFF_F + sizeof(float) * SLICE_FLOATS;        // 2 bucket tiles
constexpr size_t POFF_HASH2 = POFF_HASH + 2 * (size_t)PHH * HP;            // 2 overlap-column tiles
constexpr size_t POFF_LUT = (POFF_HASH2 + 2 * (size_t)PHH * OVW + 15) & ~(size_t)15;
constexpr size_t POFF_RING = POFF_LUT + LUT_WORDS * sizeof(unsigned);
constexpr size_t POFF_Q = POFF_RING + sizeof(float) * RING * SP;           // 2 chunks of column chains
constexpr size_t POFF_MBAR = POFF_Q + sizeof(float) * 2 * QCHUNK;          // filter-slice mbarrier
constexpr size_t POFF_INDONE = POFF_MBAR + 16;                             // split H2D: "the whole input plane has arrived" (chain leader -> filter warps)
constexpr size_t PIPE_SMEM_BYTES = POFF_INDONE + 16;
static_assert(PIPE_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(((PTH_MAX + 14) / 2 + 1) * LRP <= PHH * HP, "low-res staging fits in the HR tile");
// TMA landing zone of stage A's low-res window (raw samples): 128-byte aligned, in the HR tile behind the float staging above
constexpr size_t POFF_TMA = (POFF_HR + sizeof(float) * ((PTH_MAX + 14) / 2 + 1) * LRP + 127) & ~(size_t)127;
static_assert(POFF_TMA + 80 * 2 * TMAP_BOX_H <= POFF_HR + sizeof(float) * PHH * HP && (PTH_MAX + 14) / 2 + 1 == TMAP_BOX_H, "TMA box fits in the HR tile");

// (barrier number and thread count are IMMEDIATES: the count is part of the instruction, which is also what lets
// compute-sanitizer's synccheck see that these are partial barriers and not a __syncthreads() some threads skip)
template <int ID, int COUNT> __device__ __forceinline__ void group_sync_c() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
#define group_sync(id, count) group_sync_c<(id), (count)>()
// barrier + OR of a predicate over the group (bar.red): every thread gets the same answer
template <int ID, int COUNT> __device__ __forceinline__ bool group_or(bool v)
{
    unsigned r;
    asm volatile("{\n .reg .pred p, q;\n setp.ne.u32 p, %1, 0;\n bar.red.or.pred q, %2, %3, p;\n selp.u32 %0, 1, 0, q;\n}" : "=r"(r) : "r"((unsigned)v), "n"(ID), "n"(COUNT) : "memory");
    return r != 0u;
}
// barrier pair selected by the bucket-tile buffer (0 / 1)
template <int ID, int COUNT> __device__ __forceinline__ void group_sync_buf(int buf) { if (buf) group_sync_c<ID + 1, COUNT>(); else group_sync_c<ID, COUNT>(); }
// producer/consumer hand-off through hardware named barriers: the waiting side blocks in bar.sync (no issue slots, unlike an
// mbarrier spin), the signalling side does not wait (bar.arrive).  count = all threads of both sides.
template <int ID, int COUNT> __device__ __forceinline__ void named_arrive_c() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT> __device__ __forceinline__ void named_arrive_buf(int buf) { if (buf) named_arrive_c<ID + 1, COUNT>(); else named_arrive_c<ID, COUNT>(); }
constexpr int BAR_PROD = 1, BAR_CONS = 2, BAR_FULL = 3, BAR_EMPTY = 5, BAR_CHAIN = 7;   // FULL/EMPTY + bucket tile index (0/1)

// Packed fp32 pairs (sm_100 FMUL2 / FFMA2): two IEEE round-to-nearest operations per instruction, i.e. the same roundings as
// two scalar instructions at half the issue slots.  Stage B keeps its chains for weight columns (m, m+1) in one 64-bit pair.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// Shared-memory loads by 32-bit shared address + compile-time byte offset.  The filter stage composes its addresses as
// (per-lane constant) + (warp-uniform tile/row part) + immediate, which maps onto the LDS [R + UR + imm] addressing mode: no
// per-load address arithmetic.  volatile: never merged or hoisted across the barriers that order them with the stores.
template <int IMM> __device__ __forceinline__ float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(IMM)); return v; }
template <int IMM> __device__ __forceinline__ unsigned lds_u8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(IMM)); return v; }
template <int IMM> __device__ __forceinline__ void lds_f32x2x2(unsigned a, f32x2 &lo, f32x2 &hi)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+%3];" : "=l"(lo), "=l"(hi) : "r"(a), "n"(IMM));
}
__device__ __forceinline__ void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
template <int IMM> __device__ __forceinline__ void lds_b32x2(unsigned a, unsigned &lo, unsigned &hi)
{
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(lo), "=r"(hi) : "r"(a), "n"(IMM));
}

// compile-time loop: f(std::integral_constant<int, 0>{}) ... f(std::integral_constant<int, N - 1>{})
template <int I, int N, typename F> __device__ __forceinline__ void static_for_impl(F &f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for_impl<I + 1, N>(f);
    }
}
template <int N, typename F> __device__ __forceinline__ void static_for(F &&f) { static_for_impl<0, N>(f); }

// Sliding-window form of stage D (exact fp32 filter): see pipe_filter_pass.  Shuffle sources of its folded lane tree, packed per
// lane-in-group q -- per 4-step unit (bits 0..15: steps 0..3, bits 16..31: steps 4..7): source lanes of level 2 (3 + 3 bits) and
// level 3 (3 bits), then one bit per step "this lane keeps the odd chain at level 1".  Generated AND verified (symbolically, against
// the reference's chain and tree order) by tools/slide_model.py.
#ifndef RAISR_STAGE_D_SLIDE
#define RAISR_STAGE_D_SLIDE 1
#endif
#ifndef RAISR_STAGE_D_WALKS
#define RAISR_STAGE_D_WALKS 1
#endif
// byte offset of ring row i & (RING-1) of the producers' S ring, i = 0 .. 2 RING - 1 (stage B: window rows by uniform table look-up)
static __constant__ int c_ringoff[32] = {0 * 568, 1 * 568, 2 * 568, 3 * 568, 4 * 568, 5 * 568, 6 * 568, 7 * 568, 8 * 568, 9 * 568, 10 * 568, 11 * 568,
                                         12 * 568, 13 * 568, 14 * 568, 15 * 568, 0 * 568, 1 * 568, 2 * 568, 3 * 568, 4 * 568, 5 * 568, 6 * 568, 7 * 568,
                                         8 * 568, 9 * 568, 10 * 568, 11 * 568, 12 * 568, 13 * 568, 14 * 568, 15 * 568};
static_assert(SP * 4 == 568 && RING == 16, "c_ringoff");
static __constant__ unsigned c_slide_tbl[8] = {0x0a521452u, 0x1a3b043bu, 0x12e40ce4u, 0x168d088du, 0x15760b76u, 0x051f1b1fu, 0x0dc013c0u, 0x09a917a9u};

// Half-precision pairs of the opt-in fp16 filter stage: IEEE binary16, round to nearest even (HMUL2 / HFMA2: one rounding per
// operation, like the vmulph / vfmadd...ph of the reference's AVX512-FP16 dot product, Raisr_AVX512FP16.cpp:227-242).
__device__ __forceinline__ unsigned f2h2(float lo, float hi) { unsigned r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ unsigned hmul2(unsigned a, unsigned b) { unsigned r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned hfma2(unsigned a, unsigned b, unsigned c) { unsigned r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ void h2f2(unsigned h2, float &lo, float &hi)
{
    asm("{\n .reg .b16 l, h;\n mov.b32 {l, h}, %2;\n cvt.f32.f16 %0, l;\n cvt.f32.f16 %1, h;\n}" : "=f"(lo), "=f"(hi) : "r"(h2));
}

// dot8() of the fp16 filter stage (overlap columns, rare path): frow = the pixel's 256-byte row of half-precision coefficients in
// the same lane-permuted order; the chain pair (2q, 2q+1) runs in half2, the 16 chain sums are added in fp32 (tree8).
__device__ __forceinline__ float dot8_h(const float *sp, const char *frow, const int (&off)[8][2], int q)
{
    unsigned acc = 0u;
    const uint2 *f2 = reinterpret_cast<const uint2 *>(frow) + q;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const uint2 f = f2[n * 8];
        const unsigned p01 = f2h2(sp[off[2 * n][0]], sp[off[2 * n][1]]), p23 = f2h2(sp[off[2 * n + 1][0]], sp[off[2 * n + 1][1]]);
        acc = (n == 0) ? hmul2(p01, f.x) : hfma2(p01, f.x, acc);
        acc = hfma2(p23, f.y, acc);
    }
    float a0, a1;
    h2f2(acc, a0, a1);
    return tree8(a0, a1, q);
}

// One sample of the upscaled plane for the producer's ring (out-of-frame coordinates are clamped: such samples only
// feed pixels that are never hashed).
template <typename PixT, int UPS>
__device__ __forceinline__ float sample_S(const PassParams &p, int Y, int X)
{
    if (UPS == 1) {
        // exact 2x: even outputs weigh low-res (j-1, j) by (1,3), odd outputs (j, j+1) by (3,1); replicate border
        const int jy = Y >> 1, jx = X >> 1;
        const int ya = min(max((Y & 1) ? jy : jy - 1, 0), p.up_src_h - 1), yb = min(max((Y & 1) ? jy + 1 : jy, 0), p.up_src_h - 1);
        const int xa = min(max((X & 1) ? jx : jx - 1, 0), p.in_w - 1), xb = min(max((X & 1) ? jx + 1 : jx, 0), p.in_w - 1);
        const PixT *ra = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)ya * p.in_pitch);
        const PixT *rb = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)yb * p.in_pitch);
        const float a = (float)((unsigned)ra[xa] >> p.in_shift), b = (float)((unsigned)ra[xb] >> p.in_shift), c = (float)((unsigned)rb[xa] >> p.in_shift), d = (float)((unsigned)rb[xb] >> p.in_shift);
        const float wya = (Y & 1) ? 3.0f : 1.0f, wyb = 4.0f - wya, wxa = (X & 1) ? 3.0f : 1.0f, wxb = 4.0f - wxa;
        const float va = ffma(wya, a, fmul(wyb, c)), vb = ffma(wya, b, fmul(wyb, d));      // exact integers < 2^24
        return floorf(fmul(fadd(ffma(wxa, va, fmul(wxb, vb)), 8.0f), 0.0625f));
    }
    const int Yc = min(max(Y, 0), p.H - 1), Xc = min(max(X, 0), p.W - 1);
    return load_S<PixT>(p, Yc, Xc, UPS != 0);
}

// Bounded acquire-spin on a flag the host's copy streams write (cuStreamWriteValue32 behind a copy).  Every copy a kernel waits for
// is enqueued before the launch, so the wait ends as soon as the copy engine gets there; should it never (a copy that failed), the
// spin gives up after ~2 s of %globaltimer, raises *err (the host call then returns RNLErrorUndefined) and lets the kernel run on:
// wrong rows in a frame that is reported as failed, never a hung device.
__device__ __forceinline__ void spin_wait_flag(const unsigned *flag, unsigned seq, unsigned *err)
{
    unsigned long long t0 = 0;
    for (unsigned n = 1;; ++n) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == seq) return;
        if ((n & 1023u) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            volatile unsigned *verr = err;                                   // page-locked host word (plain stores: no PCIe atomics needed)
            const bool failed = verr && *verr != 0;                          // another CTA has already given up
            if (failed || t - t0 > 2000000000ull) {
                if (verr) { *verr = 1u; __threadfence_system(); }
                return;
            }
        }
    }
}

// same, for the input watermark (frame sequence number << 16 | rows that have arrived): wait until *flag has reached `need`, in
// wrap-around arithmetic (the sequence number wraps every 65536 frames).  Returns the value seen ("every row" after a timeout).
__device__ __forceinline__ unsigned spin_wait_flag_reach(const unsigned *flag, unsigned need, unsigned *err)
{
    unsigned long long t0 = 0;
    for (unsigned n = 1;; ++n) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - need) >= 0) return v;
        if ((n & 1023u) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            volatile unsigned *verr = err;
            const bool failed = verr && *verr != 0;
            if (failed || t - t0 > 2000000000ull) {
                if (verr) { *verr = 1u; __threadfence_system(); }
                return need | 0xffffu;                                        // "everything": no second 2 s wait in a frame that has failed
            }
        }
    }
}

// same, for a counter that only grows during the launch: wait until *flag >= need
__device__ __forceinline__ void spin_wait_flag_geq(const unsigned *flag, unsigned need, unsigned *err)
{
    unsigned long long t0 = 0;
    for (unsigned n = 1;; ++n) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= need) return;
        if ((n & 1023u) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            volatile unsigned *verr = err;
            const bool failed = verr && *verr != 0;
            if (failed || t - t0 > 2000000000ull) {
                if (verr) { *verr = 1u; __threadfence_system(); }
                return;
            }
        }
    }
}

// Split H2D: the lower part of the input plane (rows >= in_split_row) may still be on its way (own stream, flagged).  Called by
// a warp group (its leader spins, the group's named barrier publishes the result) before the first access to such rows.  Not inlined, scalar arguments only (taking the
// address of the kernel parameter block would move every access to it into local memory).
// The chunk hand-off of the two producer roles.  The chain warps and the bucket warps reach this named barrier from two different
// places of the code: legal (the .aligned rule of bar.sync is per warp), but compute-sanitizer's synccheck expects the participants
// of a barrier to meet at ONE instruction and reports "divergent threads".  -DRAISR_SYNCCHECK_BUILD routes both roles through a
// single out-of-line barrier instruction (synccheck: 0 errors, profiles/r2_sanitizer.txt); the call costs 2.5 % (0.6054 vs
// 0.5904 ms per 4K frame), so the shipped build keeps the barrier inline.
// (Round 2, measured and removed: splitting this hand-off so that a bucket warp no longer waits for the OTHER bucket warps -- chain ->
// bucket over one mbarrier per chain buffer, bucket -> chain over mbarriers (0.551 ms) or over named barriers the bucket warps only
// arrive on (0.586 ms), against 0.531 ms for the single 512-thread barrier: a waiting warp that polls takes issue slots from the warps
// it waits for, and this kernel is bound by issue slots.  bar.sync blocks for free.)
#ifdef RAISR_SYNCCHECK_BUILD
static __device__ __noinline__ void producer_chunk_barrier() { group_sync_c<BAR_PROD, NPT>(); }
#else
static __device__ __forceinline__ void producer_chunk_barrier() { group_sync_c<BAR_PROD, NPT>(); }
#endif

template <int BAR, int COUNT>
static __device__ __noinline__ bool wait_split_input(const unsigned *flag, unsigned seq, int rows_needed, int rows_total, unsigned *err, bool leader,
                                                     volatile unsigned *s_all)
{
    bool all = false;
    if (leader) {
        const unsigned v = spin_wait_flag_reach(flag, (seq << 16) | (unsigned)rows_needed, err);
        all = (int)(v - ((seq << 16) | (unsigned)rows_total)) >= 0;
        if (all) *s_all = 1u;                                                // seen by the filter warps behind this tile's FULL barrier
    }
    return group_or<BAR, COUNT>(all);                                        // the group's barrier; true: the whole plane has arrived, no further waits
}

// Chroma planes: plain cheap upscale (Raisr.cpp:1373-1388) of this CTA's share of the planes, slice sl of nslices, by the filter
// warps (NCT threads, ct = thread index in the group).  Not inlined: keeps its registers out of the filter loop's allocation.
struct ChromaParams {
    int chroma_n;
    const void *in[2]; size_t in_pitch[2]; void *out[2]; size_t out_pitch[2];
    int c_in_w, c_in_h, c_W, c_H;
    const int *c_xmap, *c_xw, *c_ymap, *c_yw;
    int c_denx, c_deny;
    int comps, shift;        // 2: one semi-planar plane (U, V interleaved);  samples carry their value in the high bits (P010: 6)
    const unsigned *chroma_ready; unsigned chroma_seq; unsigned *chroma_done; unsigned *err_flag;
};
// Work is cut over "virtual planes" vp = plane * comps + component: planar chroma has two planes of one component, semi-planar
// chroma (NV12 / P010, what NVDEC and NVENC use) one plane of two interleaved components.
template <typename PixT>
__device__ __noinline__ void chroma_slice_fn(const ChromaParams cp, int sl, int nslices, int ct)
{
    const ChromaParams &p = cp;
    const int comps = p.comps, shift = p.shift, nvp = p.chroma_n * comps;
    if (sl == 0 && p.chroma_ready) {                                 // the planes' H2D copies run on their own stream
        if (ct == 0) spin_wait_flag(p.chroma_ready, p.chroma_seq, p.err_flag);
        group_sync(BAR_CONS, NCT);
    }
    if (p.c_W == 2 * p.c_in_w && p.c_H == 2 * p.c_in_h && p.c_denx == 4 && p.c_deny == 4) {
        // exact 2x: one item = 2 output rows x 8 output columns from a 3 x 6 low-res window (replicate border = clamped
        // coordinates); even outputs weigh low-res (i-1, i) by (1,3), odd outputs (i, i+1) by (3,1): (9a+3b+3c+d+8)>>4
        const int gw8 = (p.c_W + 7) / 8, per_plane8 = gw8 * p.c_in_h;
        const long long total8 = (long long)nvp * per_plane8;
        const int c0 = (int)(total8 * blockIdx.x / gridDim.x), c1 = (int)(total8 * (blockIdx.x + 1) / gridDim.x);
        const int s0 = c0 + (int)((long long)(c1 - c0) * sl / nslices), s1 = c0 + (int)((long long)(c1 - c0) * (sl + 1) / nslices);
        for (int idx = s0 + ct; idx < s1; idx += NCT) {
            const int vp = idx / per_plane8, rem = idx - vp * per_plane8;
            const int pl = vp / comps, cc = vp - pl * comps;
            const int jb = rem / gw8, ib = rem - jb * gw8;
            const char *ibase = static_cast<const char *>(p.in[pl]);
            const PixT *rows[3] = {reinterpret_cast<const PixT *>(ibase + (size_t)max(jb - 1, 0) * p.in_pitch[pl]),
                                   reinterpret_cast<const PixT *>(ibase + (size_t)jb * p.in_pitch[pl]),
                                   reinterpret_cast<const PixT *>(ibase + (size_t)min(jb + 1, p.c_in_h - 1) * p.in_pitch[pl])};
            unsigned L[3][6];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 6; ++k) L[r][k] = (unsigned)rows[r][min(max(4 * ib - 1 + k, 0), p.c_in_w - 1) * comps + cc] >> shift;
            unsigned o[2][8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int m = e >> 1;
                unsigned h[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) h[r] = (e & 1) ? 3u * L[r][m + 1] + L[r][m + 2] : L[r][m] + 3u * L[r][m + 1];
                o[0][e] = ((h[0] + 3u * h[1] + 8u) >> 4) << shift;
                o[1][e] = ((3u * h[1] + h[2] + 8u) >> 4) << shift;
            }
            const int X = 8 * ib;
#pragma unroll
            for (int yy = 0; yy < 2; ++yy) {
                PixT *orow = reinterpret_cast<PixT *>(static_cast<char *>(p.out[pl]) + (size_t)(2 * jb + yy) * p.out_pitch[pl]) + X * comps + cc;
                const bool vec = comps == 1 && (reinterpret_cast<uintptr_t>(orow) % (8 * sizeof(PixT))) == 0 && X + 7 < p.c_W;
                if (vec && sizeof(PixT) == 1) {
                    *reinterpret_cast<uint2 *>(orow) = make_uint2(o[yy][0] | (o[yy][1] << 8) | (o[yy][2] << 16) | (o[yy][3] << 24),
                                                                 o[yy][4] | (o[yy][5] << 8) | (o[yy][6] << 16) | (o[yy][7] << 24));
                } else if (vec) {
                    *reinterpret_cast<uint4 *>(orow) = make_uint4(o[yy][0] | (o[yy][1] << 16), o[yy][2] | (o[yy][3] << 16),
                                                                 o[yy][4] | (o[yy][5] << 16), o[yy][6] | (o[yy][7] << 16));
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (X + e < p.c_W) orow[e * comps] = (PixT)o[yy][e];
                }
            }
        }
    } else {
        PassParams pc{};
        pc.in_w = p.c_in_w; pc.in_h = p.c_in_h; pc.xmap = p.c_xmap; pc.xw = p.c_xw; pc.ymap = p.c_ymap; pc.yw = p.c_yw;
        pc.denx = p.c_denx; pc.deny = p.c_deny;
        const int gw = (p.c_W + 3) / 4;
        const int per_plane = gw * p.c_H;
        const long long total = (long long)nvp * per_plane;
        const int c0 = (int)(total * blockIdx.x / gridDim.x), c1 = (int)(total * (blockIdx.x + 1) / gridDim.x);   // this CTA's groups
        const int s0 = c0 + (int)((long long)(c1 - c0) * sl / nslices), s1 = c0 + (int)((long long)(c1 - c0) * (sl + 1) / nslices);
        for (int idx = s0 + ct; idx < s1; idx += NCT) {
            const int vp = idx / per_plane, rem = idx - vp * per_plane;
            const int pl = vp / comps, cc = vp - pl * comps;
            const int Y = rem / gw, X = (rem - Y * gw) * 4;
            PixT *orow = reinterpret_cast<PixT *>(static_cast<char *>(p.out[pl]) + (size_t)Y * p.out_pitch[pl]) + X * comps + cc;
            unsigned v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (X + e < p.c_W) ? bilinear_sample<PixT>(pc, p.in[pl], p.in_pitch[pl], Y, X + e, comps, cc, shift) << shift : 0u;
            const bool vec = comps == 1 && ((reinterpret_cast<uintptr_t>(orow) % (4 * sizeof(PixT))) == 0) && X + 3 < p.c_W;
            if (vec) {
                if (sizeof(PixT) == 1) *reinterpret_cast<uint32_t *>(orow) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
                else *reinterpret_cast<uint2 *>(orow) = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (X + e < p.c_W) orow[e * comps] = (PixT)v[e];
            }
        }
    }
    if (sl == nslices - 1 && p.chroma_done) {                        // this CTA's share is written: tell the host's copy stream
        group_sync(BAR_CONS, NCT);
        if (ct == 0) {
            __threadfence();
            atomicAdd(p.chroma_done, 1u);
        }
    }
}

// ---- state a thread carries across the passes of ONE launch ------------------------------------------------------------
struct PipeCarry {
    int iter;                // tiles this CTA has started so far (bucket tile buffer = iter & 1; FULL/EMPTY hand-offs run on across passes)
    int total_iters;         // tiles this CTA walks in the whole launch (all passes)
    unsigned gk;             // producers: running chunk number (chain buffer = gk & 1)
    unsigned nload;          // filter warps: filter-slice loads issued so far (mbarrier phase)
    unsigned ntma;           // filter warps: tensor-map loads of the low-res input window issued so far (mbarrier phase)
};

// Tiles of a pass and this CTA's first one.  A launch walks the tiles of pass A and then those of pass B as ONE round-robin
// sequence T = blockIdx.x + k * gridDim.x (T < tiles(A): tile T of pass A, else tile T - tiles(A) of pass B): no bubble between
// the passes, every CTA simply keeps going.
__device__ __forceinline__ int pass_tiles(const PassParams &p)
{
    return ((p.W + TW - 1) / TW) * ((p.row1 - p.row0 + p.tile_h - 1) / p.tile_h);
}

// Input rows [lo, hi] (of the pass's input plane) a tile with output rows [y0, y0 + th) reads: the +-7 halo, through the upscale.
template <int UPS>
__device__ __forceinline__ void tile_input_rows(const PassParams &p, int y0, int th, int &lo, int &hi)
{
    const int yfirst = max(0, y0 - 8), ylast = min(p.H - 1, y0 + th + 6);
    if (UPS == 0) { lo = yfirst; hi = ylast; }
    else if (UPS == 1) { lo = max(0, (yfirst >> 1) - 1); hi = (ylast >> 1) + 1; }
    else { lo = max(0, (__ldg(p.ymap + yfirst) >> 1) - 1); hi = (__ldg(p.ymap + ylast) >> 1) + 1; }
}

// Chained second pass: its input plane is the first pass's output, produced by the SAME launch.  The first pass counts finished
// tiles per tile row (dep_done, zeroed by the host before the launch); before a group reads input rows [lo, hi] its leader waits
// until every tile row of pass A that covers them is complete (bounded spin, ld.acquire.gpu: also drops stale L1 lines), the
// group's named barrier publishes that.  `known` = tile rows [0, known) already seen complete (tiles come in row order).
template <int BAR, int COUNT>
static __device__ __noinline__ void wait_rows_done(const unsigned *done, unsigned need, int ty_hi, int known, unsigned *err, bool leader)
{
    if (leader)
        for (int ty = known; ty <= ty_hi; ++ty) spin_wait_flag_geq(done + ty, need, err);
    group_sync_c<BAR, COUNT>();
}

// =========================== producers: buckets of tile i -> bucket tile [i & 1] ===========================
// chain warps (stage B) and bucket warps (stage C) of one pass; tid = thread index inside the producer group
// The Gaussian weights come from immutable constant tables selected by the sample type (raisr_gw_tables.h: 8-bit and 10-bit; 16-bit
// samples run on the phase-sequential kernel): compile-time constant-bank addresses, i.e. uniform operands of FMUL2 in both inlined
// copies of this function.  (Read from the parameter block, the second copy of a chained launch staged all 66 weights through
// registers and spilled them inside the column-chain loop: 1.50 ms instead of 1.21 ms for the two passes at 1080p->4K.)
// FASTH: opt-in separable "fast hash" (RAISR_NUMERICS_FAST_HASH).  The reference's 2-D Gaussian is rank one up to the 6 printed digits
// of its literals (|w[i][k] - a[i] a[k]| <= 2.3e-6 w[i][k], a[i] = sqrt(w[i][i])), so GTWG = sum_i sum_k w[i][k] g g can run as an
// 11-tap vertical pass (stage B: 3 values per position instead of 18 chains) and an 11-tap horizontal pass (stage C).  ~70 instead
// of ~230 flops per pixel -- and other roundings: buckets are no longer bit-identical to the reference (measured agreement: DESIGN.md).
template <typename PixT, int PT, int UPS, bool DEP, bool FASTH, int NUMK>
__device__ __forceinline__ void pipe_producer_pass(const PassParams &p, unsigned char *smem_raw, int tile0, PipeCarry &cy, int tid)
{
    const float (&gw)[6][6] = (sizeof(PixT) == 1) ? c_gw8 : c_gw10;
    const float (&ga)[6] = (sizeof(PixT) == 1) ? c_ga8 : c_ga10;
    const unsigned *sLut = reinterpret_cast<const unsigned *>(smem_raw + POFF_LUT);
    float *sRing = reinterpret_cast<float *>(smem_raw + POFF_RING);
    float *sQ = reinterpret_cast<float *>(smem_raw + POFF_Q);
    const int th = p.tile_h, hh = th + 2;
    const int W = p.W, H = p.H;
    const int gx = (W + TW - 1) / TW;
    const int ntiles = pass_tiles(p);
    HashCtx hc{p.qstr0, p.qstr1, p.qcoh0, p.qcoh1, p.numerics, p.qangle, p.nangles, p.quarter, p.half, sLut + (tid & LUT_REPLICA_MASK) * LUT_REPLICA_WORDS, sLut + LUT_WORDS / 2 + (tid & LUT_REPLICA_MASK) * LUT_REPLICA_WORDS,
               p.lut_rsqrtps, p.lut_rcpps, (float)p.nangles};

    // 2x fast path: one thread = one 2x2 block of the upscaled plane from 4 low-res samples.  Ring row s even <-> frame row
    // Y = y0-7+s odd = 2j+1 and s+1 <-> 2j+2: both interpolate low-res rows (j, j+1) with weights (3,1) / (1,3); likewise the
    // columns sx = 2t, 2t+1.  Split into load and store so that the global-memory latency hides behind stage B.
    auto load_block = [&](int s, int t, int y0, int x0, unsigned &ab, unsigned &cd) {   // raw samples, two per register
        const int j = (y0 - 7 + s) >> 1, i = (x0 - 7 + 2 * t) >> 1;
        const int ya = min(max(j, 0), p.up_src_h - 1), yb = min(max(j + 1, 0), p.up_src_h - 1);
        const int xa = min(max(i, 0), p.in_w - 1), xb = min(max(i + 1, 0), p.in_w - 1);
        const PixT *ra = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)ya * p.in_pitch);
        const PixT *rb = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)yb * p.in_pitch);
        ab = ((unsigned)ra[xa] >> p.in_shift) | (((unsigned)ra[xb] >> p.in_shift) << 16);
        cd = ((unsigned)rb[xa] >> p.in_shift) | (((unsigned)rb[xb] >> p.in_shift) << 16);
    };
    auto store_block = [&](int s, int t, unsigned ab, unsigned cd) {
        const float a = (float)(ab & 0xffffu), b = (float)(ab >> 16), c = (float)(cd & 0xffffu), d = (float)(cd >> 16);
        const float t0 = ffma(3.0f, a, c), t1 = ffma(3.0f, b, d);     // row Y   : 3*row(j) + row(j+1)
        const float u0 = ffma(3.0f, c, a), u1 = ffma(3.0f, d, b);     // row Y+1 : row(j) + 3*row(j+1)
        float *r0 = sRing + (s & (RING - 1)) * SP + 2 * t, *r1 = sRing + ((s + 1) & (RING - 1)) * SP + 2 * t;
        r0[0] = floorf(fmul(fadd(ffma(3.0f, t0, t1), 8.0f), 0.0625f));   // col X   : 3*col(i) + col(i+1)
        r0[1] = floorf(fmul(fadd(ffma(3.0f, t1, t0), 8.0f), 0.0625f));   // col X+1 : col(i) + 3*col(i+1)
        r1[0] = floorf(fmul(fadd(ffma(3.0f, u0, u1), 8.0f), 0.0625f));
        r1[1] = floorf(fmul(fadd(ffma(3.0f, u1, u0), 8.0f), 0.0625f));
    };
    const bool chain_warp = tid >= NBT;                                  // sub-role: stage B (column chains) or stage C (buckets)
    const int lt = tid & (NBT - 1);
    // The two producer roles run as one pipeline over all chunks of all tiles of this CTA: the chain warps are one chunk
    // ahead of the bucket warps (double-buffered chains, buffer = running chunk number & 1, one BAR_PROD per chunk), also
    // across tile (and pass) boundaries, where the chain warps refill the ring while the bucket warps finish the previous tile.
    int in_known = p.in_split_row;                                       // split H2D: input rows [0, in_known) are known to have arrived
    int dep_known = 0;                                                   // DEP: tile rows [0, dep_known) of the previous pass seen complete
    for (int tile = tile0; tile < ntiles; tile += gridDim.x, ++cy.iter) {
        const int buf = cy.iter & 1;
        unsigned char *sHash = smem_raw + POFF_HASH + (size_t)buf * PHH * HP;
        unsigned char *sHash2 = smem_raw + POFF_HASH2 + (size_t)buf * PHH * OVW;
        const int ty = tile / gx, tx = tile - ty * gx;
        const int x0 = tx * TW, y0 = p.row0 + ty * th;
        const int rl = lt / QW, q = lt - rl * QW;            // this thread's row of a chunk and chain column / pixel column
        const int rlu = __shfl_sync(0xffffffffu, rl, 0);     // (the same for the whole warp: uniform for the compiler too)
        const unsigned gk = cy.gk;
        // ---- B: column chains of chunk kb, one position per thread (gradients straight from the ring) -> sQ[kb & 1].
        // Unconditional (positions outside the hashed rows produce values nobody reads): straight-line code.
        auto stage_B = [&](int kb) {
            const int s0 = RBP * kb + rl;                            // S row above the first gradient row
            if (FASTH) {
                // vertical pass of the separable form: V_k3(r, x) = sum_i a[i] (g1 g2)(r + i, x)
                f32x2 accA = 0ull;                                   // (xx, xy)
                float accY = 0.0f;                                   // yy
                float vprev = sRing[((s0) & (RING - 1)) * SP + q + 1];
                const float *row = sRing + ((s0 + 1) & (RING - 1)) * SP + q;
                float vcur = row[1];
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const float *nrow = sRing + ((s0 + 2 + i) & (RING - 1)) * SP + q;
                    const float vnext = nrow[1];
                    const float gxv = fsub(vnext, vprev), gyv = fsub(row[2], row[0]);
                    const float a = ga[i < 6 ? i : 10 - i];
                    const float ax = fmul(a, gxv), ay = fmul(a, gyv);
                    accA = fma2(pack2(ax, ax), pack2(gxv, gyv), accA);
                    accY = ffma(ay, gyv, accY);
                    vprev = vcur; vcur = vnext; row = nrow;
                }
                float *qd = sQ + ((gk + kb) & 1u) * QCHUNK + (rl * 18) * QW + q;
                float xx, xy;
                unpack2(accA, xx, xy);
                qd[0] = xx; qd[QW] = xy; qd[2 * QW] = accY;
                return;
            }
            f32x2 acc[3][3];                                         // [weight-column pair (2mm, 2mm+1)][gx*gx, gx*gy, gy*gy]
#pragma unroll
            for (int mm = 0; mm < 3; ++mm) acc[mm][0] = acc[mm][1] = acc[mm][2] = 0ull;
            // window row j = ring row (s0 + j) & (RING-1).  The row of a chunk is the same for a whole warp, so the byte offset of every
            // window row is a UNIFORM load from a constant table (LDCU, then LDS [R + UR]): one instruction per row
            const int r0u = (RBP * kb + rlu) & (RING - 1);
            const char *pq = reinterpret_cast<const char *>(sRing + q);
            auto wrow = [&](int j) { return reinterpret_cast<const float *>(pq + c_ringoff[r0u + j]); };
            float vprev = wrow(0)[1];
            const float *row = wrow(1);
            float vcur = row[1];
#pragma unroll
            for (int i = 0; i < 11; ++i) {
                const float *nrow = wrow(2 + i);
                const float vnext = nrow[1];
                const float gxv = fsub(vnext, vprev);                // GetGx (Raisr_AVX512.cpp:54-57)
                const float gyv = fsub(row[2], row[0]);              // GetGy (Raisr_AVX512.cpp:59-62)
                const f32x2 gx2 = pack2(gxv, gxv), gy2 = pack2(gyv, gyv);
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) {
                    const f32x2 w2 = pack2(gw[i < 6 ? i : 10 - i][2 * mm], gw[i < 6 ? i : 10 - i][2 * mm + 1]);   // rows i and 10-i share their weights
                    const f32x2 px = mul2(gx2, w2), py = mul2(gy2, w2);  // round(g * w), Raisr_AVX512.cpp:64-67
                    acc[mm][0] = fma2(px, gx2, acc[mm][0]);
                    acc[mm][1] = fma2(px, gy2, acc[mm][1]);
                    acc[mm][2] = fma2(py, gy2, acc[mm][2]);
                }
                vprev = vcur; vcur = vnext; row = nrow;
            }
            float *qd = sQ + ((gk + kb) & 1u) * QCHUNK + (rl * 18) * QW + q;
#pragma unroll
            for (int mm = 0; mm < 3; ++mm)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float lo, hi;
                    unpack2(acc[mm][k], lo, hi);
                    qd[(2 * mm * 3 + k) * QW] = lo;
                    qd[((2 * mm + 1) * 3 + k) * QW] = hi;
                }
        };
        // ---- C: bucket of chunk kc, one pixel per thread <- sQ[kc & 1].  The 16-wide hash runs unconditionally (garbage in,
        // nothing stored, for threads without a hashed pixel); only the row-tail columns take the 8-wide branch.
        // what depends on the thread's column only: decided once per tile, not once per chunk
        const int cj = min(q, HW - 1), cc = x0 - 1 + cj;
        const bool col_hashed = cc >= 6 && cc < p.c_end, col_tail = cc >= p.tail_start, col_both = cc < p.ov_end;
        const bool col_ov = cc >= p.tail_start && cc < p.tail_start + OVW, col_out = p.hash_out != nullptr && cj >= 1 && cj <= TW;
        auto stage_C = [&](int kc) {
            const int h = RBP * kc + rl;
            const int j = cj;
            const int r = y0 - 1 + h, c = cc;
            float g[3];
            const float *qs = sQ + ((gk + kc) & 1u) * QCHUNK + (rl * 18) * QW + j;
#pragma unroll
            for (int k3 = 0; k3 < 3; ++k3) {
                if (FASTH) {                                         // horizontal pass: sum_k a[k] V_k3(r, c + k)
                    float acc = 0.0f;
#pragma unroll
                    for (int k = 0; k < 11; ++k) acc = ffma(ga[k < 6 ? k : 10 - k], qs[k3 * QW + k], acc);
                    g[k3] = acc;
                } else {
                    float lane[11];
#pragma unroll
                    for (int k = 0; k < 11; ++k) {
                        const int m = k < 6 ? k : 10 - k;
                        lane[k] = qs[(m * 3 + k3) * QW + k];
                    }
                    g[k3] = tree_sum(lane);
                }
            }
            const bool hashed = col_hashed && r >= 6 && r < H - 6;
            int hv = hash_bucket<true, NUMK>(hc, g[0], g[1], g[2]), hv2 = 255;
            if (hashed && col_tail) {
                const int h8 = hash_bucket<false, NUMK>(hc, g[0], g[1], g[2]);
                if (col_both && hv != h8) hv2 = hv;             // also hashed by the 16-wide block before (kept if the 8-wide result is out of range)
                hv = h8;
            }
            if (!hashed) hv = 255;
            if (q < HW && h < hh) {
                if (col_out && hashed && r >= p.row0 && r < p.row1) p.hash_out[(size_t)r * W + c] = hv;
                sHash[h * HP + j] = (unsigned char)hv;
                if (col_ov) sHash2[h * OVW + (c - p.tail_start)] = (unsigned char)hv2;
            }
        };
        const int nchunks = hh / RBP;
        if (chain_warp) {
            // split H2D: the lower part of the input plane may still be on its way (own stream, flagged).  Both reader groups
            // look for themselves (the filter warps' stage A): neither is ordered behind the other's wait.
            if (!DEP && p.in_ready && in_known < p.in_h) {
                int in_lo, in_last;
                tile_input_rows<UPS>(p, y0, th, in_lo, in_last);
                if (in_last >= in_known) {                                   // the copies land in row order: wait for the watermark to pass this tile's last row
                    in_known = min(in_last + 1, p.in_h);
                    if (wait_split_input<BAR_CHAIN, NBT>(p.in_ready, p.in_seq, in_known, p.in_h, p.err_flag, lt == 0,
                                                         reinterpret_cast<volatile unsigned *>(smem_raw + POFF_INDONE)))
                        in_known = p.in_h;
                }
            }
            if (DEP) {                                                   // chained pass: the previous pass's rows this tile reads
                int in_lo, in_hi;
                tile_input_rows<UPS>(p, y0, th, in_lo, in_hi);
                const int ty_hi = min(p.dep_ny - 1, (min(in_hi, p.dep_row1 - 1) - p.dep_row0) / p.dep_th);
#ifdef RAISR_EXP_NO_DEPWAIT
                if (false) {
#else
                if (ty_hi >= dep_known) {
#endif
                    wait_rows_done<BAR_CHAIN, NBT>(p.dep_done, (unsigned)p.dep_gx, ty_hi, dep_known, p.err_flag, lt == 0);
                    dep_known = ty_hi + 1;
                }
            }
            // S ring: tile-local S row s (frame row y0-7+s) lives in ring row s & (RING-1); chunk k (filtered rows 2k, 2k+1)
            // reads rows 2k .. 2k+13.  Rows 0 .. 13 up front (the previous tile's last B is done: BAR_PROD), every chunk then
            // fetches the two rows of the NEXT chunk into the slots of rows 2k-2, 2k-1.
            if (UPS == 1) {
                for (int idx = lt; idx < ((RING - RBP) / 2) * (SW / 2); idx += NBT) {
                    const int sp2 = idx / (SW / 2), t = idx - sp2 * (SW / 2);
                    unsigned ab, cd;
                    load_block(2 * sp2, t, y0, x0, ab, cd);
                    store_block(2 * sp2, t, ab, cd);
                }
            } else {
                for (int idx = lt; idx < (RING - RBP) * SW; idx += NBT) {
                    const int s = idx / SW, sx = idx - s * SW;
                    sRing[s * SP + sx] = sample_S<PixT, UPS>(p, y0 - 7 + s, x0 - 7 + sx);
                }
            }
            group_sync(BAR_CHAIN, NBT);
            for (int k = 0; k < nchunks; ++k) {
                const bool more = k + 1 < nchunks;
                unsigned pab = 0u, pcd = 0u;
                if (UPS == 1 && more && lt < SW / 2) load_block(RBP * k + RING - RBP, lt, y0, x0, pab, pcd);
                stage_B(k);
                if (more) {
                    if (UPS == 1) {
                        if (lt < SW / 2) store_block(RBP * k + RING - RBP, lt, pab, pcd);
                    } else {
                        // (measured: fetching these samples before stage B like the 2x path costs more registers than the latency
                        // it hides -- 0.326 instead of 0.289 ms for the two 1.5x passes at 720p)
                        for (int idx = lt; idx < RBP * SW; idx += NBT) {
                            const int s = RBP * k + RING - RBP + idx / SW, sx = idx % SW;
                            sRing[(s & (RING - 1)) * SP + sx] = sample_S<PixT, UPS>(p, y0 - 7 + s, x0 - 7 + sx);
                        }
                    }
                }
                producer_chunk_barrier();                                // B(k) published; C(k-1) is done with the other chain buffer; ring rows of chunk k+1 in place
            }
        } else {
            if (cy.iter >= 2) group_sync_buf<BAR_EMPTY, NBT + NCT>(buf);    // the filter warps are done with this bucket tile (two tiles ago)
            for (int k = 0; k < nchunks; ++k) {
                producer_chunk_barrier();                                // B(k) is complete
                stage_C(k);
            }
            named_arrive_buf<BAR_FULL, NBT + NCT>(buf);                     // bucket tile [buf] is complete (bar.arrive orders this thread's writes)
        }
        cy.gk += (unsigned)nchunks;
    }
}

// =========================== filter warps: S tile, filter, blend and store of tile i ===========================
// F16: opt-in fp16 filter stage (RAISR_NUMERICS_FP16_FILTER): the filter slice holds half-precision coefficients (half the
// bytes per slice and per row), patch values are converted to half2 and the 16 chains run as HMUL2/HFMA2; the 16 chain sums are
// then added in fp32 with the same tree.  Buckets are untouched (the hash stays fp32); Y is NOT bit-identical in this mode.
template <typename PixT, int PT, int UPS, bool DEP, bool F16, int NUMK>
__device__ __forceinline__ void pipe_filter_pass(const PassParams &p, unsigned char *smem_raw, int tile0, PipeCarry &cy, int ct)
{
    float *sS = reinterpret_cast<float *>(smem_raw + POFF_S);
    float *sHR = reinterpret_cast<float *>(smem_raw + POFF_HR);
    float *sF = reinterpret_cast<float *>(smem_raw + POFF_F);
    unsigned long long *mslice = reinterpret_cast<unsigned long long *>(smem_raw + POFF_MBAR);
    const int th = p.tile_h, hh = th + 2;
    const int W = p.W, H = p.H;
    const int gx = (W + TW - 1) / TW;
    const int ntiles = pass_tiles(p);

    const int lane = ct & 31, cwarp = __shfl_sync(0xffffffffu, ct >> 5, 0);   // warp-uniform for the compiler too
    const int g = lane >> 3, q = lane & 7;
    constexpr int JS = (PT == 4) ? 2 : 1;
    constexpr int U = 4;
    constexpr int NCOLS = (HW + JS - 1) / JS;
    constexpr int NBLK = (NCOLS + 4 * U - 1) / (4 * U);
    constexpr int ULAST = (NCOLS - 4 * U * (NBLK - 1) + 3) / 4;
    constexpr int ROWB = F16 ? 256 : 512;                                 // bytes of one filter row in the slice
    constexpr int ROWSH = F16 ? 8 : 9;
    int off[8][2];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int k = 16 * m + 2 * q + e;
            off[m][e] = (k < 121) ? (k / 11) * SP + (k % 11) : 0;
        }
    const float flo = (float)p.lo, fhi = (float)p.hi;
    const int slice_bytes = p.nbuckets * ROWB;
    // per-lane constant parts of the fast block's shared addresses (bytes): pixel group g, lane q of the group
    constexpr bool SLIDE = (RAISR_STAGE_D_SLIDE != 0) && !F16;         // sliding-window form of stage D (below)
    unsigned poff[8][2];                                              // patch tap (m, e) of the group's pixel in block column 0
    if constexpr (!SLIDE) {
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                poff[m][e] = smem_u32(sS) + 4u * (unsigned)(off[m][e] + SP + 1 + g * JS);
                asm volatile("" : "+r"(poff[m][e]));                  // keep the 16 addresses in registers (no rematerialisation per block)
            }
    }
    // sliding form: the lane's own positions of the column strip of pixel group g, linearised as k0 = 11 R + c (R = strip row, c =
    // patch column): ring slot s holds the pair k0 = 16 s + 2q, 16 s + 2q + 1 (and, 16 strip rows further down, the pair s + 11)
    unsigned spk[11];                                                 // packed: low half = byte offset of the pair's first tap in the S tile, high half = second tap
    unsigned stbl = 0u, hrlane = 0u, hlane = 0u;
    if constexpr (SLIDE) {
#pragma unroll
        for (int s = 0; s < 11; ++s) {
            const int k = 16 * s + 2 * q;
            const unsigned ox = 4u * (unsigned)((k / 11) * SP + (k % 11) + SP + 1 + g * JS);
            const unsigned oy = 4u * (unsigned)(((k + 1) / 11) * SP + ((k + 1) % 11) + SP + 1 + g * JS);
            spk[s] = ox | (oy << 16);
            asm volatile("" : "+r"(spk[s]));                          // keep the 11 words in registers (no rematerialisation per item)
        }
        stbl = c_slide_tbl[q];
        hrlane = smem_u32(sHR) + 4u * (unsigned)(g * JS + 2 * (q & 3) * HP);   // HR cell of the pixel whose sum ends up in this lane (step q & 3 of a unit)
        hlane = (unsigned)(g * JS + 2 * (q & 3) * HP);                        // ... and its bucket
    }
    const unsigned fbase = smem_u32(sF) + (unsigned)(ROWB / 32) * (unsigned)q;   // this lane's 16 (fp16: 8) bytes of every filter step
    const unsigned hroff = smem_u32(sHR) + 4u * (unsigned)((g + 4 * (q & 3)) * JS);   // HR column of the pixel whose sum ends up in this lane
    // ---- chroma planes: plain cheap upscale (Raisr.cpp:1373-1388), 4 pixels per thread.  This CTA's share of the planes is
    // cut into up to three slices, one per tile from the third tile on (the planes' H2D copies have landed by then), done while the filter warps would otherwise wait for
    // the bucket tile: no launch of its own, and finished early enough for the host to copy the planes out while the luma runs on.
    const int cta_tiles = (ntiles > tile0) ? (ntiles - tile0 + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nslices = min(3, max(1, cta_tiles - 2));                   // early in the frame: the host may copy the planes out while the luma runs on
    int slices_done = 0;
    auto chroma_slice = [&](int sl) {                                 // by value: never take the address of the kernel parameter block
        ChromaParams cp;
        cp.chroma_n = p.chroma_n;
        for (int i = 0; i < 2; ++i) { cp.in[i] = p.chroma[i].in; cp.in_pitch[i] = p.chroma[i].in_pitch; cp.out[i] = p.chroma[i].out; cp.out_pitch[i] = p.chroma[i].out_pitch; }
        cp.c_in_w = p.c_in_w; cp.c_in_h = p.c_in_h; cp.c_W = p.c_W; cp.c_H = p.c_H;
        cp.c_xmap = p.c_xmap; cp.c_xw = p.c_xw; cp.c_ymap = p.c_ymap; cp.c_yw = p.c_yw; cp.c_denx = p.c_denx; cp.c_deny = p.c_deny;
        cp.comps = p.c_comps; cp.shift = p.c_shift;
        cp.chroma_ready = p.chroma_ready; cp.chroma_seq = p.chroma_seq; cp.chroma_done = p.chroma_done; cp.err_flag = p.err_flag;
        chroma_slice_fn<PixT>(cp, sl, nslices, ct);
    };
    int resident_type = -1;                                           // pixel type whose filter slice (of THIS pass's table) is in shared memory
    int local_iter = 0;
    bool input_complete = false;
    for (int tile = tile0; tile < ntiles; tile += gridDim.x, ++cy.iter, ++local_iter) {
        const int buf = cy.iter & 1;
        const unsigned char *sHash = smem_raw + POFF_HASH + (size_t)buf * PHH * HP;
        const unsigned char *sHash2 = smem_raw + POFF_HASH2 + (size_t)buf * PHH * OVW;
        const int ty = tile / gx, tx = tile - ty * gx;
        const int x0 = tx * TW, y0 = p.row0 + ty * th;
        // Stage A reads the same input rows as the chain warps' ring of this tile, and nothing else orders it behind THEIR waits
        // (split H2D flag; chained pass: rows of the previous pass).  Where such a wait can be pending the filter warps take this
        // tile's FULL barrier BEFORE stage A instead of after it: the buckets of the tile exist only once the chain warps have
        // seen the flag (leader's ld.acquire -> BAR_CHAIN -> BAR_PROD -> BAR_FULL: causality is cumulative).  The producers run a
        // tile ahead, so the barrier has usually completed long before and nothing is lost.
        bool full_taken = false;
#ifndef RAISR_EXP_NO_EARLY_FULL
        if (DEP) {
            group_sync_buf<BAR_FULL, NBT + NCT>(buf);
            full_taken = true;
        } else
#endif
        if (p.in_ready && !input_complete) {
            int in_lo, in_last;
            tile_input_rows<UPS>(p, y0, th, in_lo, in_last);
            if (in_last >= p.in_split_row) {
                group_sync_buf<BAR_FULL, NBT + NCT>(buf);
                full_taken = true;
                // "the chain leader has seen the whole plane": the word is written asynchronously to this group, so the threads may
                // read different values -- the decision must be uniform (it moves a barrier), hence the OR over the group
                input_complete = group_or<BAR_CONS, NCT>(*reinterpret_cast<volatile unsigned *>(smem_raw + POFF_INDONE) != 0u);
            }
        }

        // ---- A: S tile (the HR tile is free until the HR := S copy below: low-res staging for the 2x path; the slice buffer
        // keeps the previous tile's last filter slice) ----
        if (UPS == 1) {
            float *sL = sHR;
            const int ly0 = (y0 - 8) >> 1, lx0 = (x0 - 8) >> 1;
            const int lrh = (th + 14) / 2 + 1;
            // Interior tiles (no clamped coordinate; never the chained second pass, whose input the SAME launch writes through the
            // generic proxy): the low-res window arrives by one TMA tensor-map load into the unused upper part of the HR tile and is
            // converted from there.  Tiles on the frame border keep the clamped scalar loads (TMA fills out-of-range elements with
            // zeros, the reference replicates the border).
            constexpr int BW = TmapBox<PixT>::W;
            const bool tma_tile = !DEP && p.use_tmap && lx0 >= 0 && ly0 >= 0 && lx0 + LRW <= p.in_w && ly0 + lrh <= p.up_src_h;
            if (tma_tile) {
                PixT *stage = reinterpret_cast<PixT *>(smem_raw + POFF_TMA);
                if (ct == 0) {
                    fence_proxy_async();                                  // the HR tile was last written through the generic proxy
                    mbar_expect_tx(mslice + 1, (unsigned)(BW * TMAP_BOX_H * sizeof(PixT)));
                    tma_load_2d(stage, p.in_tmap, lx0 & ~(TmapBox<PixT>::ALIGN - 1), ly0, mslice + 1);
                }
                const int xoff = lx0 & (TmapBox<PixT>::ALIGN - 1);
                mbar_wait(mslice + 1, cy.ntma & 1u);
                ++cy.ntma;
                for (int idx = ct; idx < lrh * LRW; idx += NCT) {
                    const int ly = idx / LRW, lx = idx - ly * LRW;
                    sL[ly * LRP + lx] = (float)((unsigned)stage[ly * BW + xoff + lx] >> p.in_shift);
                }
            } else {
                for (int idx = ct; idx < lrh * LRW; idx += NCT) {
                    const int ly = idx / LRW, lx = idx - ly * LRW;
                    const int yy = min(max(ly0 + ly, 0), p.up_src_h - 1), xx = min(max(lx0 + lx, 0), p.in_w - 1);
                    sL[ly * LRP + lx] = (float)((unsigned)reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)yy * p.in_pitch)[xx] >> p.in_shift);
                }
            }
            group_sync(BAR_CONS, NCT);
            for (int idx = ct; idx < (lrh - 1) * (LRW - 1); idx += NCT) {
                const int bi = idx / (LRW - 1) + 1, bj = idx - (bi - 1) * (LRW - 1) + 1;
                const float *l = sL + bi * LRP + bj;
                const float a0 = l[-LRP - 1], a1 = l[-LRP], b0 = l[-1], b1 = l[0];
                const float vo0 = ffma(3.0f, a0, b0), vo1 = ffma(3.0f, a1, b1);
                const float ve0 = ffma(3.0f, b0, a0), ve1 = ffma(3.0f, b1, a1);
                float *d = sS + (2 * bi - 2) * SP + 2 * bj - 2;
                d[0] = floorf(fmul(fadd(ffma(3.0f, vo0, vo1), 8.0f), 0.0625f));
                d[1] = floorf(fmul(fadd(ffma(3.0f, vo1, vo0), 8.0f), 0.0625f));
                d[SP] = floorf(fmul(fadd(ffma(3.0f, ve0, ve1), 8.0f), 0.0625f));
                d[SP + 1] = floorf(fmul(fadd(ffma(3.0f, ve1, ve0), 8.0f), 0.0625f));
            }
        } else {
            for (int idx = ct; idx < (th + 14) * SW; idx += NCT) {
                const int sy = idx / SW, sx = idx - sy * SW;
                const int Y = y0 - 7 + sy, X = x0 - 7 + sx;
                float v = 0.0f;
                if (Y >= 0 && Y < H && X >= 0 && X < W) v = load_S<PixT>(p, Y, X, UPS != 0);
                sS[sy * SP + sx] = v;
            }
        }
        group_sync(BAR_CONS, NCT);
        // HR := S; the filter phase overwrites accepted pixels
        for (int idx = ct; idx < hh * HW; idx += NCT) {
            const int h = idx / HW, j = idx - h * HW;
            sHR[h * HP + j] = sS[(h + 6) * SP + j + 6];
        }
        if (p.chroma_n > 0 && local_iter >= 2 && slices_done < nslices) chroma_slice(slices_done++);
        group_sync(BAR_CONS, NCT);                                        // S / HR tile complete
        if (!full_taken) group_sync_buf<BAR_FULL, NBT + NCT>(buf);           // buckets of this tile are ready

        // ---- D: 121-tap filter, one pixel type at a time ----
        const bool has_ov = (x0 - 1 + HW > p.tail_start) && (x0 - 1 < p.tail_start + OVW);
        // pixel types in alternating order (0..PT-1, then PT-1..0): the slice the previous tile ended with is still resident
        for (int ti = 0; ti < PT; ++ti) {
            const int t = (local_iter & 1) ? PT - 1 - ti : ti;
            const bool load = t != resident_type;
            if (load && ct == 0) {
                fence_proxy_async();
                mbar_expect_tx(mslice, (unsigned)slice_bytes);
                const char *src = reinterpret_cast<const char *>(p.filters) + (size_t)t * slice_bytes;
                const int piece = slice_bytes / 4;
                for (int i = 0; i < 4; ++i) bulk_g2s(reinterpret_cast<char *>(sF) + i * piece, src + i * piece, (unsigned)piece, mslice);
            }
            const int jfirst = (PT == 4) ? ((((x0 - 1 - 5) & 1) == (t & 1)) ? 0 : 1) : 0;
            const int hfirst = (PT == 4) ? ((((y0 - 1 - 5) & 1) == (t >> 1)) ? 0 : 1) : 0;
            const int nrows = (hh - hfirst + JS - 1) / JS;
            if (load) {
                mbar_wait(mslice, cy.nload & 1u);
                ++cy.nload;
                resident_type = t;
            }
            // One warp iteration = UU groups of 4 pixels of one tile row: addresses = per-lane constant + uniform + immediate,
            // packed FMUL2/FFMA2 for the chain pair (2q, 2q+1), and the 16 -> 1 lane tree of the 4 pixel groups folded into 8 shuffles:
            // after the first exchange (t8) every lane keeps half of the pixels it holds and sends the other half, so that the
            // same additions as tree8() end up in lanes q & 3 == u (both halves q < 4 and q >= 4 hold the final sum).
            const unsigned hbase = smem_u32(sHash) + (unsigned)(g * JS);
            const unsigned lastrow = (unsigned)p.nbuckets - 1u;
            // One warp item = one tile row of this pixel type (NBLK blocks of 4 pixel groups), software-pipelined over its blocks: the
            // loads and FMAs of block b+1 are issued before the lane tree of block b (two register sets alternate), so the shuffle
            // latency of the tree hides under the next block's shared-memory loads.  The filter warps are bound by exactly that
            // stream of loads, not by issue slots or total shared-memory traffic: this alone took the 4K frame from 0.593 to 0.567 ms.
            // (Rows outside [6, H-6) are skipped; columns from c_end on carry bucket 255 and drop their result.)
            auto blk_mac = [&](auto uu, const int h, const int jc0, f32x2 (&acc)[4], unsigned (&acch)[4], unsigned (&hv)[4]) {
                constexpr int UU = decltype(uu)::value;
                const unsigned ub = 4u * (unsigned)(h * SP + jc0);
                const unsigned hva = hbase + (unsigned)(h * HP + jc0);
                unsigned fa[4];
                hv[0] = lds_u8<0>(hva); hv[1] = lds_u8<4 * JS>(hva);
                hv[2] = (UU > 2) ? lds_u8<8 * JS>(hva) : 255u; hv[3] = (UU > 3) ? lds_u8<12 * JS>(hva) : 255u;
#pragma unroll
                for (int u = 0; u < 4; ++u) fa[u] = fbase + (min(hv[u], lastrow) << ROWSH);
                auto step = [&](auto nn) {
                    constexpr int n = decltype(nn)::value;
                    const unsigned q0 = poff[2 * n][0] + ub, q1 = poff[2 * n][1] + ub, q2 = poff[2 * n + 1][0] + ub, q3 = poff[2 * n + 1][1] + ub;
                    auto one = [&](auto uc) {
                        constexpr int u = decltype(uc)::value;
                        if (u < UU) {
                            if (F16) {
                                unsigned f01, f23;
                                lds_b32x2<n * 64>(fa[u], f01, f23);
                                const unsigned p01 = f2h2(lds_f32<16 * JS * u>(q0), lds_f32<16 * JS * u>(q1));
                                const unsigned p23 = f2h2(lds_f32<16 * JS * u>(q2), lds_f32<16 * JS * u>(q3));
                                acch[u] = (n == 0) ? hmul2(p01, f01) : hfma2(p01, f01, acch[u]);
                                acch[u] = hfma2(p23, f23, acch[u]);
                            } else {
                                f32x2 fxy, fzw;
                                lds_f32x2x2<n * 128>(fa[u], fxy, fzw);
                                const f32x2 p01 = pack2(lds_f32<16 * JS * u>(q0), lds_f32<16 * JS * u>(q1));
                                const f32x2 p23 = pack2(lds_f32<16 * JS * u>(q2), lds_f32<16 * JS * u>(q3));
                                acc[u] = (n == 0) ? mul2(p01, fxy) : fma2(p01, fxy, acc[u]);
                                acc[u] = fma2(p23, fzw, acc[u]);
                            }
                        }
                    };
                    one(std::integral_constant<int, 0>{}); one(std::integral_constant<int, 1>{});
                    one(std::integral_constant<int, 2>{}); one(std::integral_constant<int, 3>{});
                };
                step(std::integral_constant<int, 0>{}); step(std::integral_constant<int, 1>{});
                step(std::integral_constant<int, 2>{}); step(std::integral_constant<int, 3>{});
            };
            auto blk_fin = [&](auto uu, const int h, const int jc0, const f32x2 (&acc)[4], const unsigned (&acch)[4], const unsigned (&hv)[4]) {
                constexpr int UU = decltype(uu)::value;
                const bool hi4 = q >= 4, b1 = (q & 2) != 0, b0 = (q & 1) != 0;
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (u < UU) {
                        float a0, a1;
                        if (F16) h2f2(acch[u], a0, a1); else unpack2(acc[u], a0, a1);
                        v[u] = fadd(hi4 ? a1 : a0, __shfl_xor_sync(0xffffffffu, hi4 ? a0 : a1, 4, 8));
                    } else v[u] = 0.0f;
                }
                const float w0 = fadd(b1 ? v[2] : v[0], __shfl_xor_sync(0xffffffffu, b1 ? v[0] : v[2], 2, 8));
                const float w1 = fadd(b1 ? v[3] : v[1], __shfl_xor_sync(0xffffffffu, b1 ? v[1] : v[3], 2, 8));
                const float x = fadd(b0 ? w1 : w0, __shfl_xor_sync(0xffffffffu, b0 ? w0 : w1, 1, 8));
                const float cur = fadd(x, __shfl_xor_sync(0xffffffffu, x, 4, 8));
                const unsigned hu = b1 ? (b0 ? hv[3] : hv[2]) : (b0 ? hv[1] : hv[0]);
                if (q < 4 && hu != 255u && cur > flo && cur < fhi)
                    sts_f32(hroff + 4u * (unsigned)(h * HP + jc0), cur);
            };
            // (Measured: carrying the pipeline on from one row of the warp to the next -- last block of a row finished under the first
            // block of the following row -- keeps both register sets live across the loop and costs more than it hides: 0.599 vs 0.567 ms.)
            // Items are row SEGMENTS of SEG = 4 blocks (4 pixel types: the whole row; one pixel type: half a row, which balances the
            // 12 warps better); the last block of a row is shorter (ULAST groups).
            constexpr int SEG = 4, NSEG = (NBLK + SEG - 1) / SEG;
            static_assert(NBLK == 4 || NBLK == 8, "row segments of 4 blocks");
            auto run_seg = [&](auto lastc, const int h, const int jbase) {
                constexpr bool LASTSEG = decltype(lastc)::value;
                f32x2 A[2][4];
                unsigned AH[2][4], HV[2][4];
                auto mac_b = [&](auto bc) {
                    constexpr int bb = decltype(bc)::value;
                    const int jc0 = jbase + bb * 4 * U * JS;
                    if constexpr (LASTSEG && bb == SEG - 1) blk_mac(std::integral_constant<int, ULAST>{}, h, jc0, A[bb & 1], AH[bb & 1], HV[bb & 1]);
                    else blk_mac(std::integral_constant<int, U>{}, h, jc0, A[bb & 1], AH[bb & 1], HV[bb & 1]);
                };
                auto fin_b = [&](auto bc) {
                    constexpr int bb = decltype(bc)::value;
                    const int jc0 = jbase + bb * 4 * U * JS;
                    if constexpr (LASTSEG && bb == SEG - 1) blk_fin(std::integral_constant<int, ULAST>{}, h, jc0, A[bb & 1], AH[bb & 1], HV[bb & 1]);
                    else blk_fin(std::integral_constant<int, U>{}, h, jc0, A[bb & 1], AH[bb & 1], HV[bb & 1]);
                };
                mac_b(std::integral_constant<int, 0>{});
                mac_b(std::integral_constant<int, 1>{}); fin_b(std::integral_constant<int, 0>{});
                mac_b(std::integral_constant<int, 2>{}); fin_b(std::integral_constant<int, 1>{});
                mac_b(std::integral_constant<int, 3>{}); fin_b(std::integral_constant<int, 2>{});
                fin_b(std::integral_constant<int, 3>{});
            };
            if constexpr (!SLIDE) {
            for (int it = cwarp; it < nrows * NSEG; it += NCW) {
                const int ri = it / NSEG, sg = it - ri * NSEG;
                const int h = hfirst + ri * JS, r = y0 - 1 + h;
                const int jbase = jfirst + sg * SEG * 4 * U * JS;
                if (r < 6 || r >= H - 6 || x0 - 1 + jbase >= p.c_end) continue;
                if (sg == NSEG - 1) run_seg(std::true_type{}, h, jbase);
                else if constexpr (NSEG > 1) run_seg(std::false_type{}, h, jbase);
            }
            } else {
            // ---- sliding-window form (the default of the exact fp32 filter) ----
            // A pixel group (8 lanes) walks DOWN a column of same-type pixels, two tile rows per step, and keeps its patch values in
            // registers.  Linearise the column strip as k0 = 11 R + c: the taps of the pixel of step N are k0 in [22 N, 22 N + 121), tap
            // k = k0 - 22 N.  Lane q owns the positions k0 = 2q, 2q + 1 (mod 16) for the whole walk; at step N these are exactly the
            // taps of the chain pair p = (q - 3N) & 7, in chain order: the ownership of the reference's 16 chains (Raisr_AVX512.cpp:
            // 134-149: chain j = taps 16 m + j, m = 0..7, then the 16-lane tree) ROTATES through the lanes instead of the values moving.
            // Per step a lane loads the 1.4 new tap pairs that enter the strip window (ring of 11 pairs = the period of 8 steps = 16
            // strip rows) instead of all 16 taps: 2.75 instead of 16 patch loads per pixel and lane.  The chain of lanes q < 3N mod 8
            // starts one pair later than that of the others; both alignments are accumulated (two independent FFMA2 chains against
            // the same coefficient registers) and the lane selects its own.  The lane tree is the same folded tree as before, with
            // shuffle sources from c_slide_tbl (the chain pairs sit in rotated lanes).  Schedule and tree are verified symbolically
            // by tools/slide_model.py; results are bit-identical to the block form above (same products, same order).
            // Positions past the strip of the last pixel (pairs a lane loads but does not own taps in) meet coefficient zero.
            constexpr int NCB = (NCOLS + 3) / 4;                          // column blocks: 4 pixel groups side by side
            constexpr int NWALK = (PT == 4) ? 1 : 2;                      // one pixel type: rows of both parities, two walks per column
            const int hs0 = (PT == 4) ? hfirst : 0;
            const int nseg = ((hh - hs0 + 1) / 2 + 7) / 8;                // walks are cut into items of 8 steps (two units of 4)
            const unsigned lastrow_s = (unsigned)p.nbuckets - 1u;
            const unsigned fbm = fbase - 128u;
            // A walk = nunits units (4 steps each) of one column block, starting at tile row h0; the 8-step cycle below repeats with the
            // ring carried across cycle boundaries (CONT: step 7 also requests what step 0 of the next cycle needs -- pair 18 = pair 7
            // of the next cycle, the coefficient units and the bucket of its first two steps -- before the bases move on by 16 rows).
            auto slide_walk = [&](const int h0, const int jc0, int nunits) {
                unsigned ubv = smem_u32(sS) + 4u * (unsigned)(h0 * SP + jc0);       // strip origin of this walk (uniform, but wanted in a vector register:
                asm volatile("" : "+r"(ubv));                                      //  one add per patch load instead of a uniform-to-vector move plus a multiply-add)
                unsigned hva = smem_u32(sHash) + (unsigned)(g * JS + h0 * HP + jc0);
                unsigned hul = smem_u32(sHash) + hlane + (unsigned)(h0 * HP + jc0);
                unsigned hra = hrlane + 4u * (unsigned)(h0 * HP + jc0);
                float Wx[11], Wy[11];
                f32x2 C[16];                                                      // coefficient units: step n uses C[8 (n & 1) + 0..7], the other half is being loaded
                float v[4];
                auto ldw = [&](auto sc, auto immc) {
                    constexpr int sl = decltype(sc)::value, imm = decltype(immc)::value;
                    unsigned lo;                                                    // (volatile: extracted per load, not hoisted into 11 more registers)
                    asm volatile("and.b32 %0, %1, 0xffff;" : "=r"(lo) : "r"(spk[sl]));
                    Wx[sl] = lds_f32<imm>(lo + ubv);
                    Wy[sl] = lds_f32<imm>((spk[sl] >> 16) + ubv);
                };
                auto ldc = [&](auto nc, const unsigned fa) {                        // the 4 coefficient units of step n: chain pair (q - t_n) & 7 of row fa
                    constexpr int n = decltype(nc)::value, tn = (3 * n) % 8, cb0 = 8 * (n & 1);
                    lds_f32x2x2<128 + 0 - 16 * tn>(fa, C[cb0 + 0], C[cb0 + 1]);
                    lds_f32x2x2<128 + 128 - 16 * tn>(fa, C[cb0 + 2], C[cb0 + 3]);
                    lds_f32x2x2<128 + 256 - 16 * tn>(fa, C[cb0 + 4], C[cb0 + 5]);
                    lds_f32x2x2<128 + 384 - 16 * tn>(fa, C[cb0 + 6], C[cb0 + 7]);
                };
                unsigned hvc = lds_u8<0>(hva), hvn = lds_u8<2 * HP>(hva);           // buckets of steps 0 and 1
                static_for<8>([&](auto sc) { ldw(sc, std::integral_constant<int, 0>{}); });
                ldc(std::integral_constant<int, 0>{}, fbm + (min(hvc, lastrow_s) << 9));
                auto cycle = [&](auto twoc, auto contc) {
                    constexpr bool TWO = decltype(twoc)::value, CONT = decltype(contc)::value;
                    constexpr int NSTEP = TWO ? 8 : 4;
                    static_assert(TWO || !CONT, "a walk continues after full cycles only");
                    static_for<NSTEP>([&](auto nc) {
                        constexpr int n = decltype(nc)::value;
                        constexpr int B = 11 * n / 8, t = (3 * n) % 8, u = n & 3, cb0 = 8 * (n & 1);
                        constexpr bool more = n + 1 < NSTEP || CONT;
                        constexpr int tn = (3 * (n + 1)) % 8;
                        unsigned hu = 255u;
                        if constexpr (more) {
                            // everything step n + 1 needs is requested before the arithmetic of step n: the pairs entering its window (their ring
                            // slots are dead at step n), its coefficient units (other half of C), and the bucket of step n + 2
                            constexpr int Lc = (n == 0) ? 7 : B + 8, Ln = (n == 7) ? 18 : 11 * (n + 1) / 8 + 8;
                            static_for<Ln - Lc>([&](auto ic) {
                                constexpr int i = Lc + 1 + decltype(ic)::value;
                                ldw(std::integral_constant<int, i % 11>{}, std::integral_constant<int, (i >= 11) ? 16 * SP * 4 : 0>{});
                            });
                            ldc(std::integral_constant<int, (n + 1) & 7>{}, fbm + (min(hvn, lastrow_s) << 9) + ((q < tn) ? 128u : 0u));
                            if constexpr (n + 2 < NSTEP || CONT) hvn = lds_u8<2 * HP * (n + 2)>(hva);
                        }
                        if constexpr (u == 3) hu = lds_u8<2 * HP * (n - 3)>(hul);   // bucket of the pixel this lane ends up with
                        f32x2 aU = 0ull, aF = 0ull;
                        static_for<8>([&](auto mc) {
                            constexpr int m = decltype(mc)::value;
                            const f32x2 wu = pack2(Wx[(B + m) % 11], Wy[(B + m) % 11]);
                            aU = (m == 0) ? mul2(wu, C[cb0 + m]) : fma2(wu, C[cb0 + m], aU);
                            if constexpr (t != 0) {
                                const f32x2 wf = pack2(Wx[(B + 1 + m) % 11], Wy[(B + 1 + m) % 11]);
                                aF = (m == 0) ? mul2(wf, C[cb0 + m]) : fma2(wf, C[cb0 + m], aF);
                            }
                        });
                        float a0, a1;
                        if constexpr (t != 0) unpack2((q < t) ? aF : aU, a0, a1); else unpack2(aU, a0, a1);
                        const bool hi = ((stbl >> (16 * (n >> 2) + 9 + u)) & 1u) != 0u;
                        v[u] = fadd(hi ? a1 : a0, __shfl_xor_sync(0xffffffffu, hi ? a0 : a1, 4, 8));
                        if constexpr (u == 3) {
                            constexpr int SH0 = 16 * (n >> 2), n0 = n - 3;
                            const bool b1 = (q & 2) != 0, b0 = (q & 1) != 0;
                            const float w0 = fadd(b1 ? v[2] : v[0], __shfl_sync(0xffffffffu, b1 ? v[0] : v[2], (int)(stbl >> SH0), 8));
                            const float w1 = fadd(b1 ? v[3] : v[1], __shfl_sync(0xffffffffu, b1 ? v[1] : v[3], (int)(stbl >> (SH0 + 3)), 8));
                            const float x = fadd(b0 ? w1 : w0, __shfl_sync(0xffffffffu, b0 ? w0 : w1, (int)(stbl >> (SH0 + 6)), 8));
                            const float cur = fadd(x, __shfl_xor_sync(0xffffffffu, x, 4, 8));
                            if (q < 4 && hu != 255u && cur > flo && cur < fhi) sts_f32(hra + 4u * (unsigned)(2 * n0 * HP), cur);
                        }
                    });
                };
                while (nunits > 2) {
                    cycle(std::true_type{}, std::true_type{});
                    ubv += 16u * SP * 4u; hva += 16u * HP; hul += 16u * HP; hra += 4u * 16u * HP;
                    nunits -= 2;
                }
                if (nunits == 2) cycle(std::true_type{}, std::false_type{});
                else cycle(std::false_type{}, std::false_type{});
            };
#if RAISR_STAGE_D_WALKS
            // Work = units of 4 steps in column-major order; every filter warp takes one contiguous range (7 or 8 of the 90 units of a
            // pixel type at 4K), i.e. one or two whole columns' worth: a window start-up per column crossing instead of one per 8 steps.
            constexpr int NCI = NCB * NWALK;
            const int nunit = ((hh - hs0 + 1) / 2 + 3) / 4;                   // units per walk
            const int nu_total = NCI * nunit;
            const int u_begin = nu_total * cwarp / NCW, u_end = nu_total * (cwarp + 1) / NCW;
            const unsigned inv_nunit = (65536u + (unsigned)nunit - 1u) / (unsigned)nunit;   // uu / nunit == (uu * inv) >> 16 for uu < 2^14 / nunit ... (uu <= 2 * NCB * 6)
            for (int uu = u_begin; uu < u_end;) {
                const int cw = (int)(((unsigned)uu * inv_nunit) >> 16), ua = uu - cw * nunit;
                const int cnt = min(nunit - ua, u_end - uu);
                const int wk = (NWALK == 2) ? (cw & 1) : 0, cb = (NWALK == 2) ? (cw >> 1) : cw;
                const int h0 = hs0 + wk + 8 * ua, jc0 = jfirst + 4 * cb * JS;
                if (x0 - 1 + jc0 < p.c_end) slide_walk(h0, jc0, cnt);
                uu += cnt;
            }
#else
            // items of (at most) two units in segment-major order: the divisor of the index split is a compile-time constant
            constexpr int NCI = NCB * NWALK;
            for (int it = cwarp; it < NCI * nseg; it += NCW) {
                const int sg = it / NCI, cw = it - sg * NCI;
                const int wk = (NWALK == 2) ? (cw & 1) : 0, cb = (NWALK == 2) ? (cw >> 1) : cw;
                const int hsw = hs0 + wk;
                const int left = (hh - hsw + 1) / 2 - 8 * sg;             // steps left in this walk
                const int h0 = hsw + 16 * sg, jc0 = jfirst + 4 * cb * JS;
                if (left <= 0 || x0 - 1 + jc0 >= p.c_end) continue;
                slide_walk(h0, jc0, left > 4 ? 2 : 1);
            }
#endif
            }
            // Columns hashed by both the 16-wide and the 8-wide variant (Raisr.cpp:1246-1250): the pass above used the 8-wide
            // bucket (the later evaluation); where that result was out of range the reference keeps the 16-wide evaluation.
            if (has_ov && p.blending == 2) {
                const int jlo = max(jfirst, p.tail_start - (x0 - 1)), jhi = min(HW, p.tail_start + OVW - (x0 - 1));
                const int j0 = jlo + ((jlo - jfirst) % JS != 0 ? JS - (jlo - jfirst) % JS : 0);   // first column of this type in the overlap
                const int ncol = (jhi > j0) ? (jhi - j0 + JS - 1) / JS : 0;
                const int total = nrows * ncol;
                for (int base = cwarp * 4; base < total; base += NCW * 4) {
                    const int pi = min(base + g, total - 1);
                    const int ri = pi / ncol, ci = pi - ri * ncol;
                    const int h = hfirst + ri * JS, j = j0 + ci * JS;
                    const int hv = sHash[h * HP + j];
                    const int hv2 = sHash2[h * OVW + (x0 - 1 + j - p.tail_start)];
                    const float *sp = sS + (h + 1) * SP + j + 1;
                    const char *fb = reinterpret_cast<const char *>(sF);
                    const float cur8 = F16 ? dot8_h(sp, fb + (hv == 255 ? 0 : hv) * ROWB, off, q) : dot8(sp, sF + (hv == 255 ? 0 : hv) * 128, off, q);
                    const float cur16 = F16 ? dot8_h(sp, fb + (hv2 == 255 ? 0 : hv2) * ROWB, off, q) : dot8(sp, sF + (hv2 == 255 ? 0 : hv2) * 128, off, q);
                    const bool ok8 = cur8 > flo && cur8 < fhi;
                    if (q == 0 && base + g < total && hv != 255 && hv2 != 255 && !ok8 && cur16 > flo && cur16 < fhi) sHR[h * HP + j] = cur16;
                }
            }
            group_sync(BAR_CONS, NCT);
        }
        // ---- E: blend + store ----
        stage_blend_store<PixT, NUMK>(p, sS, sHR, sHash, x0, y0, th, ct, NCT);
        group_sync(BAR_CONS, NCT);                                        // S / HR are rewritten by the next tile's stage A
        if (cy.iter + 2 < cy.total_iters) named_arrive_buf<BAR_EMPTY, NBT + NCT>(buf);   // bucket tile may be refilled (two tiles on, possibly in the next pass)
        if (ct == 0 && (p.band_done || p.rows_done)) {
            __threadfence();                                              // the tile's stores (ordered before by the barrier) -> device scope
            if (p.band_done && !(p.out_tail && y0 >= p.tail_row0)) atomicAdd(p.band_done + ty / p.band_tiles_y, 1u);
            if (p.rows_done) atomicAdd(p.rows_done + ty, 1u);             // chained launch: the next pass waits for whole tile rows
        }
    }
    if (p.chroma_n > 0)
        while (slices_done < nslices) chroma_slice(slices_done++);       // CTAs with fewer than three tiles
}

// UPSA / UPSB: upscale flavour of the first / second pass of the launch (0 = none, 1 = exact 2x, 2 = axis maps); UPSB = -1: one
// pass.  With two passes (pb.dep_done set by the host, cooperative launch: all CTAs resident) the second pass reads the plane
// the first one writes; tile rows are handed over through pa.rows_done / pb.dep_done.
// NV: numerics variant of the launch -- 0: exact (bit-identical to the reference; IEEE or X86 numerics decided at run time), 1: fp16
// filter stage, 2: separable fast hash, 4: exact with the X86 numerics compiled in (the default on a library with the tables: the hot
// loops carry no IEEE code).
template <typename PixT, int PT, int UPSA, int UPSB, int NV>
__global__ void __launch_bounds__(NTP, 1) raisr_frame_pipe_kernel(const __grid_constant__ PassParams pa, const __grid_constant__ PassParams pb)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned *sLut = reinterpret_cast<unsigned *>(smem_raw + POFF_LUT);
    unsigned long long *mslice = reinterpret_cast<unsigned long long *>(smem_raw + POFF_MBAR);
    const int tid0 = threadIdx.x;

    if (pa.numerics != 0) lut14_fill(sLut, pa.lut_rsqrt14, pa.lut_rcp14, tid0, NTP);
    for (int i = tid0; i < 2 * PHH * (HP - HW); i += NTP)                 // pad columns of both bucket tiles: "not hashed"
        smem_raw[POFF_HASH + (size_t)(i / (HP - HW)) * HP + HW + i % (HP - HW)] = 255;
    if (tid0 == 0) {
        *reinterpret_cast<volatile unsigned *>(smem_raw + POFF_INDONE) = 0u;
        mbar_init(mslice, 1);
        mbar_init(mslice + 1, 1);                                           // tensor-map loads of stage A
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // this CTA's share of the launch: tiles blockIdx.x, +gridDim.x, ... of the combined sequence (pass A, then pass B)
    const int G = (int)gridDim.x, b = (int)blockIdx.x;
    const int tA = pass_tiles(pa);
    const int tB = (UPSB >= 0) ? pass_tiles(pb) : 0;
    const int nA = (tA > b) ? (tA - b + G - 1) / G : 0;
    const int tileB0 = b + nA * G - tA;                                   // first tile of pass B for this CTA (>= 0)
    const int nB = (tB > tileB0) ? (tB - tileB0 + G - 1) / G : 0;
    PipeCarry cy{0, nA + nB, 0u, 0u, 0u};

    // warps [0, NCW) are the filter warps, warps [NCW, NCW + NPW) the producers (bucket warps, then chain warps)
    if (tid0 >= NCT) {
        const int tid = tid0 - NCT;
        // hand registers to the filter warpgroups; the chain warps (18 accumulators) need fewer than the hash
        if (tid >= NBT) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CHAIN_REGS));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BUCKET_REGS));
        pipe_producer_pass<PixT, PT, UPSA, false, NV == 2, (NV == 4 ? 1 : -1)>(pa, smem_raw, b, cy, tid);
        if constexpr (UPSB >= 0) pipe_producer_pass<PixT, PT, UPSB, true, NV == 2, (NV == 4 ? 1 : -1)>(pb, smem_raw, tileB0, cy, tid);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONS_REGS));
        pipe_filter_pass<PixT, PT, UPSA, false, NV == 1, (NV == 4 ? 1 : -1)>(pa, smem_raw, b, cy, tid0);
        if constexpr (UPSB >= 0) pipe_filter_pass<PixT, PT, UPSB, true, NV == 1, (NV == 4 ? 1 : -1)>(pb, smem_raw, tileB0, cy, tid0);
    }
}

}  // namespace raisr
