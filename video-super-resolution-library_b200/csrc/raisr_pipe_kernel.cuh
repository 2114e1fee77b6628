// raisr_pipe_kernel.cuh -- persistent, warp-specialised form of the RAISR pass kernel.
//
// Same arithmetic as raisr_pass_kernel (raisr_kernels.cuh: stages A-E, bit-identical results); what changes is the
// schedule.  The profile of the phase-sequential kernel shows two kinds of stages: B/C (structure tensor + bucket) are
// bound by FP32/ALU issue with the shared-memory pipe idle, D (121-tap filter out of the shared filter slice) is bound
// by the shared-memory pipe with the FMA pipe idle.  Here one CTA per SM loops over tiles with two warp groups:
//   producer warps : tile i+1 -- stream the upscaled rows through a 16-row ring, column chains, buckets  -> bucket tile[(i+1)&1]
//   consumer warps : tile i   -- S tile, per-type filter slices (cp.async.bulk), 8-lane filter, blend, store
// so the FMA-bound and the LSU-bound work of neighbouring tiles overlap on the same SM.  The hand-off is a pair of bucket
// tiles guarded by hardware named barriers (bar.arrive / bar.sync, full and empty per tile buffer: the waiting side blocks
// without consuming issue slots), group-local synchronisation uses two more named barriers.
//
// Register file: the CTA is launched with 80 registers per thread; the producer warpgroups drop to 48 and the consumer
// warpgroups rise to 96 (setmaxnreg).  Producer: the upscaled rows of the next chunk are fetched from global memory before
// stage B and stored into free ring slots after stage C (latency hidden); stage B uses packed FMUL2/FFMA2.
//
// STATUS: bit-identical to raisr_pass_kernel (same GPU parity suite), and the default: 8 producer + 16 consumer warps.
#pragma once
#include "raisr_kernels.cuh"

namespace raisr {

constexpr int NTP = 768;                     // threads per CTA of the pipelined kernel: 8 producer + 16 consumer warps
constexpr int NPW = 8;                       // producer warps (2 warpgroups)
constexpr int NCW = NTP / 32 - NPW;          // consumer warps (3 warpgroups)
constexpr int NPT = NPW * 32, NCT = NCW * 32;
constexpr int PROD_REGS = 56, CONS_REGS = 88;    // setmaxnreg targets: 256*48 + 512*96 <= 768*80 registers of the CTA
constexpr int RBP = 2;                       // filtered rows per producer chunk (RBP * QW == NPT positions)
constexpr int RING = 16;                     // rows of the producer's S ring (>= RBP + 12, power of two)
static_assert(RBP * QW == NPT && RING >= 2 * RBP + 12 && (RING & (RING - 1)) == 0 && RBP % 2 == 0 && SW % 2 == 0, "producer geometry");

constexpr size_t POFF_S = 0;
constexpr size_t POFF_HR = POFF_S + sizeof(float) * SH * SP;
constexpr size_t POFF_F = (POFF_HR + sizeof(float) * HH * HP + 127) & ~(size_t)127;
constexpr size_t POFF_HASH = POFF_F + sizeof(float) * SLICE_FLOATS;        // 2 bucket tiles
constexpr size_t POFF_HASH2 = POFF_HASH + 2 * (size_t)HH * HP;             // 2 overlap-column tiles
constexpr size_t POFF_LUT = (POFF_HASH2 + 2 * (size_t)HH * OVW + 15) & ~(size_t)15;
constexpr size_t POFF_RING = POFF_LUT + 256 * sizeof(uint2);
constexpr size_t POFF_Q = POFF_RING + sizeof(float) * RING * SP;
constexpr size_t POFF_MBAR = POFF_Q + sizeof(float) * RBP * 18 * QW;       // filter-slice mbarrier
constexpr size_t PIPE_SMEM_BYTES = POFF_MBAR + 16;
static_assert(PIPE_SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void group_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// producer/consumer hand-off through hardware named barriers: the waiting side blocks in bar.sync (no issue slots, unlike an
// mbarrier spin), the signalling side does not wait (bar.arrive).  count = all threads of both sides.
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
constexpr int BAR_PROD = 1, BAR_CONS = 2, BAR_FULL = 3, BAR_EMPTY = 5;   // FULL/EMPTY + bucket tile index (0/1)

// Packed fp32 pairs (sm_100 FMUL2 / FFMA2): two IEEE round-to-nearest operations per instruction, i.e. the same roundings as
// two scalar instructions at half the issue slots.  Stage B keeps its chains for weight columns (m, m+1) in one 64-bit pair.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

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
        const float a = (float)ra[xa], b = (float)ra[xb], c = (float)rb[xa], d = (float)rb[xb];
        const float wya = (Y & 1) ? 3.0f : 1.0f, wyb = 4.0f - wya, wxa = (X & 1) ? 3.0f : 1.0f, wxb = 4.0f - wxa;
        const float va = ffma(wya, a, fmul(wyb, c)), vb = ffma(wya, b, fmul(wyb, d));      // exact integers < 2^24
        return floorf(fmul(fadd(ffma(wxa, va, fmul(wxb, vb)), 8.0f), 0.0625f));
    }
    const int Yc = min(max(Y, 0), p.H - 1), Xc = min(max(X, 0), p.W - 1);
    return load_S<PixT>(p, Yc, Xc, UPS != 0);
}

template <typename PixT, int PT, int UPS>
__global__ void __launch_bounds__(NTP, 1) raisr_pass_pipe_kernel(const PassParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sS = reinterpret_cast<float *>(smem_raw + POFF_S);
    float *sHR = reinterpret_cast<float *>(smem_raw + POFF_HR);
    float *sF = reinterpret_cast<float *>(smem_raw + POFF_F);
    uint2 *sLut = reinterpret_cast<uint2 *>(smem_raw + POFF_LUT);
    float *sRing = reinterpret_cast<float *>(smem_raw + POFF_RING);
    float *sQ = reinterpret_cast<float *>(smem_raw + POFF_Q);
    unsigned long long *mbars = reinterpret_cast<unsigned long long *>(smem_raw + POFF_MBAR);
    unsigned long long *mslice = mbars;

    const int tid0 = threadIdx.x;
    const int th = p.tile_h, hh = th + 2;
    const int W = p.W, H = p.H;
    const int gx = (W + TW - 1) / TW;
    const int ntiles = gx * ((p.row1 - p.row0 + th - 1) / th);

    if (p.numerics != 0 && tid0 < 256) sLut[tid0] = (tid0 < 128) ? p.lut_rsqrt14[tid0] : p.lut_rcp14[tid0 - 128];
    if (tid0 == 0) {
        mbar_init(mslice, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    HashCtx hc{p.qstr0, p.qstr1, p.qcoh0, p.qcoh1, p.numerics, p.qangle, p.nangles, p.quarter, p.half, sLut, sLut + 128, p.lut_rsqrtps, p.lut_rcpps};

    // warps [0, NCW) are consumers, warps [NCW, NCW + NPW) producers (PIPE_PROD_FIRST: the other way round)
#ifdef PIPE_PROD_FIRST
    const bool is_prod = tid0 < NPT;
    const int tid = tid0, ct = tid0 - NPT;
#else
    const bool is_prod = tid0 >= NCT;
    const int tid = tid0 - NCT, ct = tid0;
#endif
    if (is_prod) {
        // =========================== producer: buckets of tile i -> bucket tile [i & 1] ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PROD_REGS));   // hand registers to the consumer warpgroups
        // 2x fast path: one thread = one 2x2 block of the upscaled plane from 4 low-res samples.  Ring row s even <-> frame row
        // Y = y0-7+s odd = 2j+1 and s+1 <-> 2j+2: both interpolate low-res rows (j, j+1) with weights (3,1) / (1,3); likewise the
        // columns sx = 2t, 2t+1.  Split into load and store so that the global-memory latency hides behind stage B.
        auto load_block = [&](int s, int t, int y0, int x0, unsigned &ab, unsigned &cd) {   // raw samples, two per register
            const int j = (y0 - 7 + s) >> 1, i = (x0 - 7 + 2 * t) >> 1;
            const int ya = min(max(j, 0), p.up_src_h - 1), yb = min(max(j + 1, 0), p.up_src_h - 1);
            const int xa = min(max(i, 0), p.in_w - 1), xb = min(max(i + 1, 0), p.in_w - 1);
            const PixT *ra = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)ya * p.in_pitch);
            const PixT *rb = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)yb * p.in_pitch);
            ab = (unsigned)ra[xa] | ((unsigned)ra[xb] << 16);
            cd = (unsigned)rb[xa] | ((unsigned)rb[xb] << 16);
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
        int iter = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
            const int buf = iter & 1;
            unsigned char *sHash = smem_raw + POFF_HASH + (size_t)buf * HH * HP;
            unsigned char *sHash2 = smem_raw + POFF_HASH2 + (size_t)buf * HH * OVW;
            const int ty = tile / gx, tx = tile - ty * gx;
            const int x0 = tx * TW, y0 = p.row0 + ty * th;
            if (iter >= 2) group_sync(BAR_EMPTY + buf, NTP);                 // the consumer is done with this bucket tile (tile i-2)
            const bool cols_hashed = (x0 - 1 + HW > 6) && (x0 - 1 < p.c_end);
            // S ring: tile-local S row s (frame row y0-7+s) lives in ring row s & (RING-1); chunk h0 reads rows h0 .. h0+RBP+11.
            // Rows 0 .. RBP+11 up front; every chunk then fetches the RBP rows of the NEXT chunk (slots the current one does not read).
            if (UPS == 1) {
                for (int idx = tid; idx < ((RBP + 12) / 2) * (SW / 2); idx += NPT) {
                    const int sp2 = idx / (SW / 2), t = idx - sp2 * (SW / 2);
                    unsigned ab, cd;
                    load_block(2 * sp2, t, y0, x0, ab, cd);
                    store_block(2 * sp2, t, ab, cd);
                }
            } else {
                for (int idx = tid; idx < (RBP + 12) * SW; idx += NPT) {
                    const int s = idx / SW, sx = idx - s * SW;
                    sRing[(s & (RING - 1)) * SP + sx] = sample_S<PixT, UPS>(p, y0 - 7 + s, x0 - 7 + sx);
                }
            }
            for (int h0 = 0; h0 < hh; h0 += RBP) {
                const int rfirst = y0 - 1 + h0;
                const bool more = h0 + RBP < hh;
                unsigned pab = 0u, pcd = 0u;
                if (UPS == 1 && more && tid < SW / 2) load_block(h0 + RBP + 12, tid, y0, x0, pab, pcd);
                group_sync(BAR_PROD, NPT);                                   // ring rows of this chunk are in place; C(previous chunk) is done with sQ
#ifdef PIPE_DBG_SKIP_BC
                const bool any_hashed = false;
#else
                const bool any_hashed = cols_hashed && (rfirst + RBP > 6) && (rfirst < H - 6);
#endif
                if (any_hashed) {
                    // ---- B: column chains, one position per thread (gradients straight from the ring) ----
                    const int rl = tid / QW, q = tid - rl * QW;
                    const int r = rfirst + rl;
                    if (r >= 6 && r < H - 6 && h0 + rl < hh) {
                        const int s0 = h0 + rl;                                  // S row above the first gradient row
                        f32x2 acc[3][3];                                         // [weight-column pair (2mm, 2mm+1)][gx*gx, gx*gy, gy*gy]
#pragma unroll
                        for (int mm = 0; mm < 3; ++mm) acc[mm][0] = acc[mm][1] = acc[mm][2] = 0ull;
                        float vprev = sRing[((s0) & (RING - 1)) * SP + q + 1];
                        const float *row = sRing + ((s0 + 1) & (RING - 1)) * SP + q;
                        float vcur = row[1];
#pragma unroll
                        for (int i = 0; i < 11; ++i) {
                            const float *nrow = sRing + ((s0 + 2 + i) & (RING - 1)) * SP + q;
                            const float vnext = nrow[1];
                            const float gxv = fsub(vnext, vprev);                // GetGx (Raisr_AVX512.cpp:54-57)
                            const float gyv = fsub(row[2], row[0]);              // GetGy (Raisr_AVX512.cpp:59-62)
                            const f32x2 gx2 = pack2(gxv, gxv), gy2 = pack2(gyv, gyv);
#pragma unroll
                            for (int mm = 0; mm < 3; ++mm) {
                                const f32x2 w2 = pack2(c_gw[i][2 * mm], c_gw[i][2 * mm + 1]);
                                const f32x2 px = mul2(gx2, w2), py = mul2(gy2, w2);  // round(g * w), Raisr_AVX512.cpp:64-67
                                acc[mm][0] = fma2(px, gx2, acc[mm][0]);
                                acc[mm][1] = fma2(px, gy2, acc[mm][1]);
                                acc[mm][2] = fma2(py, gy2, acc[mm][2]);
                            }
                            vprev = vcur; vcur = vnext; row = nrow;
                        }
                        float *qd = sQ + (rl * 18) * QW + q;
#pragma unroll
                        for (int mm = 0; mm < 3; ++mm)
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                float lo, hi;
                                unpack2(acc[mm][k], lo, hi);
                                qd[(2 * mm * 3 + k) * QW] = lo;
                                qd[((2 * mm + 1) * 3 + k) * QW] = hi;
                            }
                    }
                    group_sync(BAR_PROD, NPT);
                }
                // ---- C: bucket, one pixel per thread ----
                {
                    const int rl = tid / QW, j = tid - rl * QW;
                    const int h = h0 + rl;
                    if (j < HW && h < hh) {
                        const int r = rfirst + rl, c = x0 - 1 + j;
                        int hv = 255, hv2 = 255;
                        if (any_hashed && r >= 6 && r < H - 6 && c >= 6 && c < p.c_end) {
                            float g[3];
                            const float *qs = sQ + (rl * 18) * QW + j;
#pragma unroll
                            for (int k3 = 0; k3 < 3; ++k3) {
                                float lane[11];
#pragma unroll
                                for (int k = 0; k < 11; ++k) {
                                    const int m = k < 6 ? k : 10 - k;
                                    lane[k] = qs[(m * 3 + k3) * QW + k];
                                }
                                g[k3] = tree_sum(lane);
                            }
                            if (c < p.tail_start) {
                                hv = hash_bucket<true>(hc, g[0], g[1], g[2]);
                            } else {
                                hv = hash_bucket<false>(hc, g[0], g[1], g[2]);
                                if (c < p.ov_end) {
                                    const int h16 = hash_bucket<true>(hc, g[0], g[1], g[2]);
                                    if (h16 != hv) hv2 = h16;
                                }
                            }
                            if (p.hash_out && r >= p.row0 && r < p.row1 && j >= 1 && j <= TW) p.hash_out[(size_t)r * W + c] = hv;
                        }
                        sHash[h * HP + j] = (unsigned char)hv;
                        if (c >= p.tail_start && c < p.tail_start + OVW) sHash2[h * OVW + (c - p.tail_start)] = (unsigned char)hv2;
                    }
                }
                // ring rows of the next chunk (slots this chunk's stage B does not read; the barrier at the top of the next chunk publishes them)
                if (more) {
                    if (UPS == 1) {
                        if (tid < SW / 2) store_block(h0 + RBP + 12, tid, pab, pcd);
                    } else {
                        for (int idx = tid; idx < RBP * SW; idx += NPT) {
                            const int s = h0 + RBP + 12 + idx / SW, sx = idx % SW;
                            sRing[(s & (RING - 1)) * SP + sx] = sample_S<PixT, UPS>(p, y0 - 7 + s, x0 - 7 + sx);
                        }
                    }
                }
            }
            named_arrive(BAR_FULL + buf, NTP);                               // bucket tile [buf] is complete (bar.arrive orders this thread's writes)
        }
    } else {
        // =========================== consumer: filter + blend of tile i ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONS_REGS));
        const int lane = ct & 31, cwarp = ct >> 5;
        const int g = lane >> 3, q = lane & 7;
        constexpr int JS = (PT == 4) ? 2 : 1;
        constexpr int U = 4;
        constexpr int NCOLS = (HW + JS - 1) / JS;
        constexpr int NBLK = (NCOLS + 4 * U - 1) / (4 * U);
        constexpr int ULAST = (NCOLS - 4 * U * (NBLK - 1) + 3) / 4;
        int off[8][2];
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 16 * m + 2 * q + e;
                off[m][e] = (k < 121) ? (k / 11) * SP + (k % 11) : 0;
            }
        const float flo = (float)p.lo, fhi = (float)p.hi;
        const int slice_bytes = p.nbuckets * 128 * (int)sizeof(float);
        const float4 *sF4 = reinterpret_cast<const float4 *>(sF) + q;
        unsigned nload = 0;
        int iter = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
            const int buf = iter & 1;
            const unsigned char *sHash = smem_raw + POFF_HASH + (size_t)buf * HH * HP;
            const unsigned char *sHash2 = smem_raw + POFF_HASH2 + (size_t)buf * HH * OVW;
            const int ty = tile / gx, tx = tile - ty * gx;
            const int x0 = tx * TW, y0 = p.row0 + ty * th;

            // ---- A: S tile (the slice buffer is free until the first slice load: low-res staging for the 2x path) ----
            if (UPS == 1) {
                float *sL = sF;
                const int ly0 = (y0 - 8) >> 1, lx0 = (x0 - 8) >> 1;
                const int lrh = (th + 14) / 2 + 1;
                for (int idx = ct; idx < lrh * LRW; idx += NCT) {
                    const int ly = idx / LRW, lx = idx - ly * LRW;
                    const int yy = min(max(ly0 + ly, 0), p.up_src_h - 1), xx = min(max(lx0 + lx, 0), p.in_w - 1);
                    sL[ly * LRP + lx] = (float)reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)yy * p.in_pitch)[xx];
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
            group_sync(BAR_FULL + buf, NTP);                                  // buckets of this tile are ready (also publishes S / HR to the group)

            // ---- D: 121-tap filter, one pixel type at a time ----
            const bool has_ov = (x0 - 1 + HW > p.tail_start) && (x0 - 1 < p.tail_start + OVW);
#ifdef PIPE_DBG_SKIP_D
            for (int t = 0; t < 0; ++t) {
#else
            for (int t = 0; t < PT; ++t) {
#endif
                if (ct == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(mslice, (unsigned)slice_bytes);
                    const char *src = reinterpret_cast<const char *>(p.filters) + (size_t)t * slice_bytes;
                    const int piece = slice_bytes / 4;
                    for (int i = 0; i < 4; ++i) bulk_g2s(reinterpret_cast<char *>(sF) + i * piece, src + i * piece, (unsigned)piece, mslice);
                }
                const int jfirst = (PT == 4) ? ((((x0 - 1 - 5) & 1) == (t & 1)) ? 0 : 1) : 0;
                const int hfirst = (PT == 4) ? ((((y0 - 1 - 5) & 1) == (t >> 1)) ? 0 : 1) : 0;
                const int nrows = (hh - hfirst + JS - 1) / JS;
                mbar_wait(mslice, nload & 1u);
                ++nload;
                auto block = [&](auto uu, const int h, const int jb) {
                    constexpr int UU = decltype(uu)::value;
                    const float *sp = sS + (h + 1) * SP + jb + 1;
                    const unsigned char *hp = sHash + h * HP + jb;
                    int hv[UU];
#pragma unroll
                    for (int u = 0; u < UU; ++u) hv[u] = (jb + 4 * JS * u < HW) ? hp[4 * JS * u] : 255;
                    float a0[UU], a1[UU];
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const float *q0 = sp + off[2 * n][0], *q1 = sp + off[2 * n][1], *q2 = sp + off[2 * n + 1][0], *q3 = sp + off[2 * n + 1][1];
#pragma unroll
                        for (int u = 0; u < UU; ++u) {
                            const float4 f = sF4[(hv[u] == 255 ? 0 : hv[u]) * 32 + n * 8];
                            const float p0 = q0[4 * JS * u], p1 = q1[4 * JS * u], p2 = q2[4 * JS * u], p3 = q3[4 * JS * u];
                            if (n == 0) { a0[u] = fmul(p0, f.x); a1[u] = fmul(p1, f.y); }
                            else { a0[u] = ffma(p0, f.x, a0[u]); a1[u] = ffma(p1, f.y, a1[u]); }
                            a0[u] = ffma(p2, f.z, a0[u]);
                            a1[u] = ffma(p3, f.w, a1[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UU; ++u) {
                        const float cur = tree8(a0[u], a1[u], q);
                        bool ok = (cur > flo) && (cur < fhi);                 // strict range test, Raisr.cpp:1192-1196
                        float res = cur;
                        const int j = jb + 4 * JS * u;
                        if (has_ov) {
                            const int c = x0 - 1 + j;
                            const int hv2 = (j < HW && c >= p.tail_start && c < p.tail_start + OVW) ? sHash2[h * OVW + (c - p.tail_start)] : 255;
                            const bool need2 = (hv2 != 255) && !ok && p.blending == 2;
                            if (__any_sync(0xffffffffu, need2)) {
                                const float cur16 = dot8(sp + 4 * JS * u, sF + (hv2 == 255 ? 0 : hv2) * 128, off, q);
                                if (need2 && cur16 > flo && cur16 < fhi) { ok = true; res = cur16; }
                            }
                        }
                        if (q == 0 && hv[u] != 255 && ok) sHR[h * HP + j] = res;
                    }
                };
                // whole rounds of (row, block) items, one item per warp; the items of the last, partial round are split into
                // single pixel groups over all warps so that no warp waits a whole item at the barrier
                const int nitems = nrows * NBLK, nfull = (nitems / NCW) * NCW;
                for (int it = cwarp; it < nfull; it += NCW) {
                    const int ri = it / NBLK, bi = it - ri * NBLK;
                    const int h = hfirst + ri * JS;
                    const int jb = jfirst + (bi * 4 * U + g) * JS;
                    if (bi < NBLK - 1) block(std::integral_constant<int, U>{}, h, jb);
                    else block(std::integral_constant<int, ULAST>{}, h, jb);
                }
                for (int rq = cwarp; rq < (nitems - nfull) * U; rq += NCW) {
                    const int it = nfull + rq / U, u = rq - (rq / U) * U;
                    const int ri = it / NBLK, bi = it - ri * NBLK;
                    if (bi == NBLK - 1 && u >= ULAST) continue;
                    block(std::integral_constant<int, 1>{}, hfirst + ri * JS, jfirst + (bi * 4 * U + 4 * u + g) * JS);
                }
                group_sync(BAR_CONS, NCT);
            }
            // ---- E: blend + store ----
            stage_blend_store<PixT>(p, sS, sHR, sHash, x0, y0, th, ct, NCT);
            group_sync(BAR_CONS, NCT);                                        // S / HR are rewritten by the next tile's stage A
            if (tile + 2 * (int)gridDim.x < ntiles) named_arrive(BAR_EMPTY + buf, NTP);   // bucket tile may be refilled (tile i+2)
            if (p.band_done && ct == 0) {
                __threadfence();
                atomicAdd(p.band_done + ty / p.band_tiles_y, 1u);
            }
        }
    }
}

}  // namespace raisr
