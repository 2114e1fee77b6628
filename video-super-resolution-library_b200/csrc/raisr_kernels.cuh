// raisr_kernels.cuh -- sm_100a device code of the RAISR luma pass and the chroma resize.
//
// One fused kernel per pass: a CTA owns a TW x TH tile of the pass's output plane and runs, entirely out
// of shared memory,
//   A  cheap upscale (exact-rational bilinear) or plain load of the integer input -> S tile (+7 halo)
//   B  structure-tensor column chains                                              -> Q chunk
//   C  per pixel: 11-lane tree sums -> eigen-analysis -> bucket; 121-tap filter    -> HR tile (+1 halo)
//   D  3x3 census blend, round, clamp, store
// Arithmetic follows the reference's fp32 AVX-512 path rounding for rounding (file:line citations are
// relative to /root/reference/Library); the organisation of the work does not.
//
// Why "column chains" (stage B).  The reference accumulates, per pixel and per patch column k, a chain
// over the 11 patch rows  acc_k = fma(round(g1*w[i][k]), g2, acc_k)  (Raisr_AVX512.cpp:64-67,104-114) and then
// tree-sums the 16 lanes (Raisr_AVX512.cpp:37-44).  A chain depends only on (row r, image column x, weight
// column k), and w[.][k] == w[.][10-k], so the chain of pixel c at k and of pixel c' = c + 2k - 10 at 10-k
// are the SAME sequence of roundings.  Computing every (r, x, min(k,10-k)) chain once is bit-identical to the
// reference and costs 6 instead of 11 chains per pixel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace raisr {

struct PassParams {
    const void *in;          // integer input plane (u8/u16); LR-sized when upscale != 0
    size_t in_pitch;         // bytes
    int in_w, in_h;
    void *out;               // integer output plane, W x H
    size_t out_pitch;        // bytes
    int W, H;
    int row0, row1;          // output rows this launch produces: [row0, row1)
    int upscale;             // 0: S = in,  1: S = resize(in)
    const int *xmap;         // [W] (i0 << 1 | step) : left source column, whether the right tap is i0+1
    const int *xw;           // [W] numerator of the right tap's weight over denx
    const int *ymap, *yw;    // same for rows, over deny
    int denx, deny;
    const float *filters;    // [216][ptypes][128]
    int ptypes;              // 4 or 1
    float qstr0, qstr1, qcoh0, qcoh1;
    int lo, hi;              // colour range
    int c_end;               // hashed columns are [6, c_end)                  (Raisr.cpp:1065-1066)
    int tail_start;          // columns >= tail_start are hashed by the 8-wide variant (Raisr.cpp:1246-1250)
    int ov_end;              // columns in [tail_start, ov_end) are hashed by BOTH variants, 8-wide last
    int numerics;            // RAISR_NUMERICS_*
    float qangle;            // angle bins / PI
    int nangles;
    int *hash_out;           // optional [H][W] bucket plane (parity tests), -1 = not hashed
    int blending;            // 2 = CountOfBitsChanged
    const uint16_t *lut_rsqrt14, *lut_rcp14, *lut_rsqrtps, *lut_rcpps;   // x86 numerics tables (may be null)
};

// Gaussian weights, folded: c_gw[i][m] = w[i][m] = w[i][10-m], m = 0..5   (Raisr_globals.h:208-264)
__constant__ float c_gw[11][6];

// ---- tile geometry -------------------------------------------------------------------------------
constexpr int NT = 256;          // threads per CTA
constexpr int TW = 116;          // output tile width  (TW + 12 == 128 chain columns)
constexpr int TH = 62;           // output tile height (TH + 2  == 64 filtered rows)
constexpr int RB = 4;            // filtered rows per chunk
constexpr int QW = TW + 12;      // chain columns per row
constexpr int HW = TW + 2;       // filtered (HR) columns per row
constexpr int SW = TW + 14;      // S tile columns
constexpr int SP = SW + 2;       // S tile pitch (floats)
constexpr int SH = TH + 14;      // S tile rows
constexpr int HP = HW + 2;       // HR tile pitch
constexpr int HH = TH + 2;       // HR tile rows
constexpr int GR = RB + 10;      // gradient rows per chunk
constexpr size_t SMEM_BYTES = sizeof(float) * ((size_t)SH * SP + (size_t)HH * HP + 2u * GR * QW + (size_t)RB * 18 * QW);
static_assert(QW == 128 && HH % RB == 0, "tile geometry");

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// ---- stage A: one sample of the cheap upscale (oracle/ipp_standin/ipp.h semantics) -----------------
template <typename PixT>
__device__ __forceinline__ float load_S(const PassParams &p, int Y, int X)
{
    if (!p.upscale) {
        const PixT *row = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)Y * p.in_pitch);
        return (float)row[X];
    }
    const int xm = __ldg(p.xmap + X), ym = __ldg(p.ymap + Y);
    const int x0 = xm >> 1, x1 = x0 + (xm & 1), y0 = ym >> 1, y1 = y0 + (ym & 1);
    const int wx1 = __ldg(p.xw + X), wy1 = __ldg(p.yw + Y);
    const int wx0 = p.denx - wx1, wy0 = p.deny - wy1;
    const PixT *ra = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)y0 * p.in_pitch);
    const PixT *rb = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)y1 * p.in_pitch);
    const unsigned a = ra[x0], b = ra[x1], c = rb[x0], d = rb[x1];
    const unsigned long long DD = (unsigned long long)p.denx * (unsigned long long)p.deny;
    if (DD * 65535ull < 0x7fffffffull) {      // small denominators (2x: 16, 1.5x: 36): 32-bit arithmetic
        const unsigned s = (unsigned)wy0 * ((unsigned)wx0 * a + (unsigned)wx1 * b) + (unsigned)wy1 * ((unsigned)wx0 * c + (unsigned)wx1 * d);
        const unsigned dd = (unsigned)DD;
        return (float)((s + dd / 2) / dd);
    }
    const unsigned long long s = (unsigned long long)wy0 * ((unsigned long long)wx0 * a + (unsigned long long)wx1 * b) +
                                 (unsigned long long)wy1 * ((unsigned long long)wx0 * c + (unsigned long long)wx1 * d);
    return (float)((s + DD / 2) / DD);
}

// ---- x86 approximation instructions via tables (numerics == X86) ------------------------------------
// vrsqrt14ps / vrcp14ps / rsqrtps / rcpps restricted to positive normal inputs, plus the special values the
// hash can produce (0, negative, NaN).  Tables: see csrc/x86_tables.h for the layout.
__device__ __forceinline__ float lut_rsqrt(const uint16_t *t, int idx_bits, float x)
{
    if (!(x >= 0.0f)) return __int_as_float(0x7fc00000) * ((x != x) ? 1.0f : -1.0f);   // NaN (x86 returns -NaN for negatives; sign of NaN never matters here)
    if (x == 0.0f) return __int_as_float(0x7f800000);
    if (x == __int_as_float(0x7f800000)) return 0.0f;
    const unsigned u = __float_as_uint(x);
    const int e = (int)(u >> 23) - 127;                 // unbiased exponent
    const unsigned par = e & 1;                          // odd exponent -> second half of the table
    const unsigned idx = (par << idx_bits) | ((u & 0x7fffffu) >> (23 - idx_bits));
    const unsigned ent = t[idx];                         // top 16 bits of the result's mantissa field for exponent slot below
    // result = 2^(-(e - par)/2) * r, r in (0.5, 1] for par = 0 -> [1/sqrt2 ...]; the table stores the full
    // 16 significant mantissa bits plus one bit telling whether the result's exponent is one lower.
    const int half = (e - (int)par) / 2;                 // exact: e - par is even (floor for negatives handled by parity)
    const unsigned man = (ent & 0x7fffu) << 8;           // 15 stored fraction bits -> mantissa bits 22..8
    const int eadj = (ent >> 15) ? -1 : 0;               // result in [0.5,1) * 2^-half  vs exactly 1.0 * 2^-half
    const int re = 127 - half + eadj;
    return __uint_as_float(((unsigned)re << 23) | man);
}

__device__ __forceinline__ float lut_rcp(const uint16_t *t, int idx_bits, float x)
{
    if (x != x) return x;
    const unsigned u = __float_as_uint(x);
    const unsigned sign = u & 0x80000000u;
    const unsigned au = u & 0x7fffffffu;
    if (au == 0) return __uint_as_float(sign | 0x7f800000u);
    if (au == 0x7f800000u) return __uint_as_float(sign);
    const int e = (int)(au >> 23) - 127;
    const unsigned idx = (au & 0x7fffffu) >> (23 - idx_bits);
    const unsigned ent = t[idx];
    const unsigned man = (ent & 0x7fffu) << 8;
    const int eadj = (ent >> 15) ? -1 : 0;               // 1/m for m in (1,2) lies in (0.5,1): exponent -1; m == 1 -> 1.0
    const int re = 127 - e + eadj;
    if (re <= 0) return __uint_as_float(sign);           // would be denormal: never reached by the hash's value range
    return __uint_as_float(sign | ((unsigned)re << 23) | man);
}

struct HashCtx {
    float qstr0, qstr1, qcoh0, qcoh1;
    int numerics;
    float qangle;
    int nangles;
    const uint16_t *rsqrt14, *rcp14, *rsqrtps, *rcpps;
};

template <bool WIDE16>
__device__ __forceinline__ float hash_sqrt(const HashCtx &h, float x)
{
    if (h.numerics == 0) return __fsqrt_rn(x);
    if (WIDE16) return lut_rcp(h.rcp14, 16, lut_rsqrt(h.rsqrt14, 15, x));     // Raisr_AVX512.cpp:200,221-222
    return lut_rcp(h.rcpps, 12, lut_rsqrt(h.rsqrtps, 12, x));                 // Raisr_AVX256.cpp:419,441-442
}

// atan2 approximation, Raisr_AVX512.cpp:151-173 (== Raisr_AVX256.cpp:366-391)
__device__ __forceinline__ float atan2_approx(float y, float x)
{
    const float ONEQTR_PI = 0.78539816339744830962f, THRQTR_PI = 2.35619449019234492885f;
    const float ay = fadd(fabsf(y), 1e-10f);
    const bool neg = x < 0.0f;
    const float num = neg ? fadd(x, ay) : fsub(x, ay);
    const float den = neg ? fsub(ay, x) : fadd(x, ay);
    const float q = __fdiv_rn(num, den);
    const float base = neg ? THRQTR_PI : ONEQTR_PI;
    const float v = ffma(ffma(fmul(0.1963f, q), q, -0.9817f), q, base);
    return (y < 0.0f) ? fmul(-1.0f, v) : v;
}

// Bucket of one pixel from its structure tensor (a, b, d).
// WIDE16: GetHashValue_AVX512_32f_16Elements (Raisr_AVX512.cpp:175-258); else the 8-wide AVX2 variant the
// AVX-512 build runs on row tails (Raisr_AVX256.cpp:393-472; Raisr.cpp:1133-1134).
template <bool WIDE16>
__device__ __forceinline__ int hash_bucket(const HashCtx &h, float a, float b, float d)
{
    const float PI_F = 3.141592653f;                             // Raisr_globals.h:29
    const float T = fadd(a, d);
    const float D = fsub(fmul(a, d), fmul(b, b));
    const float s = hash_sqrt<WIDE16>(h, fsub(fmul(fmul(T, T), 0.25f), D));
    const float hT = fmul(T, 0.5f);
    const float L1 = fadd(hT, s), L2 = fsub(hT, s);
    const float x = (b != 0.0f) ? fsub(L1, d) : 1.0f;
    float ang = atan2_approx(b, x);
    ang = fadd(ang, (ang < 0.0f) ? PI_F : 0.0f);
    const float s1 = hash_sqrt<WIDE16>(h, L1), s2 = hash_sqrt<WIDE16>(h, L2);
    const float coh = __fdiv_rn(fsub(s1, s2), fadd(fadd(s1, s2), 0.00000000000000001f));
    const float fa = floorf(fmul(ang, h.qangle));
    const int ai = (fa >= 0.0f) ? ((fa < (float)h.nangles) ? (int)fa : h.nangles - 1) : 0;   // NaN -> INT_MIN -> max(.,0) = 0
    int si, ci;
    if (WIDE16) {
        si = (h.qstr0 <= L1) + (h.qstr1 <= L1);
        ci = (h.qcoh0 <= coh) + (h.qcoh1 <= coh);
    } else {
        si = 2 - ((L1 <= h.qstr0) + (L1 <= h.qstr1));
        ci = 2 - ((coh <= h.qcoh0) + (coh <= h.qcoh1));
    }
    return ai * 9 + si * 3 + ci;
}

// 11 lane values -> the reference's 16-lane tree (Raisr_AVX512.cpp:37-44).  Pixel A of a pair occupies
// lanes 1..11, pixel B lanes 2..12 (Raisr_AVX512.cpp:107-114), hence two pairings.
__device__ __forceinline__ float tree_sum_A(const float *k)
{
    const float t40 = fadd(k[7], k[3]);
    const float t41 = fadd(fadd(k[0], k[8]), k[4]);
    const float t42 = fadd(fadd(k[1], k[9]), k[5]);
    const float t43 = fadd(fadd(k[2], k[10]), k[6]);
    return fadd(fadd(t40, t42), fadd(t41, t43));
}
__device__ __forceinline__ float tree_sum_B(const float *k)
{
    const float t40 = fadd(k[6], fadd(k[2], k[10]));
    const float t41 = fadd(k[7], k[3]);
    const float t42 = fadd(fadd(k[0], k[8]), k[4]);
    const float t43 = fadd(fadd(k[1], k[9]), k[5]);
    return fadd(fadd(t40, t42), fadd(t41, t43));
}

// 121-tap filter in the reference's order: 16 lane chains over 8 chunks, then the tree
// (DotProdPatch_AVX512_32f, Raisr_AVX512.cpp:134-149).  sp = &S[r-5][c-5] in the shared tile.
__device__ __forceinline__ float dot_patch(const float *sp, const float *__restrict__ f)
{
    float acc[16];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        float fv[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(f + 16 * m + 4 * q));
            fv[4 * q] = v.x; fv[4 * q + 1] = v.y; fv[4 * q + 2] = v.z; fv[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = 16 * m + j;
            if (k < 121) {
                const float pv = sp[(k / 11) * SP + (k % 11)];
                acc[j] = (m == 0) ? fmul(pv, fv[j]) : ffma(pv, fv[j], acc[j]);
            }
        }
    }
    float t8[8], t4[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) t8[j] = fadd(acc[j], acc[j + 8]);
#pragma unroll
    for (int j = 0; j < 4; ++j) t4[j] = fadd(t8[j], t8[j + 4]);
    return fadd(fadd(t4[0], t4[2]), fadd(t4[1], t4[3]));
}

template <typename PixT>
__global__ void __launch_bounds__(NT, 1) raisr_pass_kernel(const PassParams p)
{
    extern __shared__ float smem[];
    float *sS = smem;                       // [SH][SP]   S rows  y0-7 .. y0+TH+6, cols x0-7 .. x0+TW+6
    float *sHR = sS + SH * SP;              // [HH][HP]   HR rows y0-1 .. y0+TH,   cols x0-1 .. x0+TW
    float *sGX = sHR + HH * HP;             // [GR][QW]
    float *sGY = sGX + GR * QW;             // [GR][QW]
    float *sQ = sGY + GR * QW;              // [RB][18][QW]

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW;
    const int y0 = p.row0 + blockIdx.y * TH;
    const int W = p.W, H = p.H;

    // ---- A: S tile ---------------------------------------------------------------------------------
    for (int idx = tid; idx < SH * SW; idx += NT) {
        const int sy = idx / SW, sx = idx - sy * SW;
        const int Y = y0 - 7 + sy, X = x0 - 7 + sx;
        float v = 0.0f;
        if (Y >= 0 && Y < H && X >= 0 && X < W) v = load_S<PixT>(p, Y, X);
        sS[sy * SP + sx] = v;
    }
    __syncthreads();

    HashCtx hc{p.qstr0, p.qstr1, p.qcoh0, p.qcoh1, p.numerics, p.qangle, p.nangles, p.lut_rsqrt14, p.lut_rcp14, p.lut_rsqrtps, p.lut_rcpps};
    const float flo = (float)p.lo, fhi = (float)p.hi;

    for (int ch = 0; ch < HH / RB; ++ch) {
        const int h0 = ch * RB;                               // first HR-tile row of the chunk
        const int rfirst = y0 - 1 + h0;                       // its frame row
        // rows of this chunk that are hashed at all (uniform per CTA)
        const bool any_hashed = (rfirst + RB > 6) && (rfirst < H - 6) && (x0 - 1 + HW > 6) && (x0 - 1 < p.c_end);
        if (any_hashed) {
            // gradients for S rows h0+1 .. h0+RB+10, chain columns q <-> S col q+1
            for (int idx = tid; idx < GR * QW; idx += NT) {
                const int g = idx / QW, q = idx - g * QW;
                const int srow = h0 + 1 + g;
                const float *s = sS + srow * SP + q + 1;
                sGX[idx] = fsub(s[SP], s[-SP]);               // GetGx: next row - previous row (Raisr_AVX512.cpp:54-57)
                sGY[idx] = fsub(s[1], s[-1]);                 // GetGy: right - left            (Raisr_AVX512.cpp:59-62)
            }
            __syncthreads();
            // ---- B: column chains ---------------------------------------------------------------------
            for (int it = tid; it < RB * QW; it += NT) {
                const int rl = it / QW, q = it - rl * QW;
                float acc[6][3];
#pragma unroll
                for (int m = 0; m < 6; ++m) acc[m][0] = acc[m][1] = acc[m][2] = 0.0f;
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const float gx = sGX[(rl + i) * QW + q], gy = sGY[(rl + i) * QW + q];
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        const float w = c_gw[i][m];
                        const float px = fmul(gx, w), py = fmul(gy, w);
                        acc[m][0] = ffma(px, gx, acc[m][0]);
                        acc[m][1] = ffma(px, gy, acc[m][1]);
                        acc[m][2] = ffma(py, gy, acc[m][2]);
                    }
                }
                float *qd = sQ + (rl * 18) * QW + q;
#pragma unroll
                for (int m = 0; m < 6; ++m)
#pragma unroll
                    for (int k = 0; k < 3; ++k) qd[(m * 3 + k) * QW] = acc[m][k];
            }
            __syncthreads();
        }
        // ---- C: bucket + filter ---------------------------------------------------------------------
        for (int it = tid; it < RB * QW; it += NT) {
            const int rl = it / QW, j = it - rl * QW;
            if (j >= HW) continue;
            const int r = rfirst + rl, c = x0 - 1 + j;
            const int h = h0 + rl;
            const float sc = sS[(h + 6) * SP + j + 6];
            float hr = sc;
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
                    g[k3] = (c & 1) ? tree_sum_B(lane) : tree_sum_A(lane);
                }
                const int pt = (p.ptypes == 4) ? ((((r - 5) & 1) << 1) | ((c - 5) & 1)) : 0;     // Raisr.cpp:1068-1096
                const float *sp = sS + (h + 1) * SP + j + 1;
                int hv;
                if (c < p.tail_start) {
                    hv = hash_bucket<true>(hc, g[0], g[1], g[2]);
                    const float cur = dot_patch(sp, p.filters + ((size_t)hv * p.ptypes + pt) * 128);
                    if (cur > flo && cur < fhi) hr = cur;                                        // Raisr.cpp:1192-1196
                } else {
                    hv = hash_bucket<false>(hc, g[0], g[1], g[2]);
                    if (c < p.ov_end) {       // first pass of the overlap: 16-wide hash (kept if the 8-wide result is invalid)
                        const int hv16 = hash_bucket<true>(hc, g[0], g[1], g[2]);
                        if (hv16 != hv) {
                            const float cur16 = dot_patch(sp, p.filters + ((size_t)hv16 * p.ptypes + pt) * 128);
                            if (cur16 > flo && cur16 < fhi) hr = cur16;
                        }
                    }
                    const float cur = dot_patch(sp, p.filters + ((size_t)hv * p.ptypes + pt) * 128);
                    if (cur > flo && cur < fhi) hr = cur;
                }
                if (p.hash_out && r >= p.row0 && r < p.row1 && j >= 1 && j <= TW) p.hash_out[(size_t)r * W + c] = hv;
            }
            sHR[h * HP + j] = hr;
        }
        __syncthreads();
    }

    // ---- D: census blend + store (CTCountOfBitsChangedSegment_AVX256_32f, Raisr_AVX256.cpp:68-166) ----
    for (int idx = tid; idx < TH * TW; idx += NT) {
        const int ty = idx / TW, tx = idx - ty * TW;
        const int Y = y0 + ty, X = x0 + tx;
        if (Y >= p.row1 || Y >= H || X >= W) continue;
        const float *s = sS + (ty + 7) * SP + tx + 7;
        const float *hq = sHR + (ty + 1) * HP + tx + 1;
        const float lc = s[0], hcv = hq[0];
        int iv;
        if (Y == 0 || X == 0 || Y == H - 1 || X == W - 1) {
            iv = (int)lc;                                           // 1-px frame: the integer upscale itself (Raisr.cpp:999-1028,1252-1265)
        } else {
            int ham = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    if (dy == 0 && dx == 0) continue;
                    ham += ((s[dy * SP + dx] < lc) != (hq[dy * HP + dx] < hcv));
                }
            const float w = fmul((float)ham, 0.125f);
            const float v = fadd(fadd(fmul(w, lc), fmul(fsub(1.0f, w), hcv)), 0.5f);
            iv = (int)floorf(v);
            iv = min(max(iv, p.lo), p.hi);
        }
        PixT *orow = reinterpret_cast<PixT *>(static_cast<char *>(p.out) + (size_t)Y * p.out_pitch);
        orow[X] = (PixT)iv;
    }
}

// ---- chroma: plain cheap upscale (Raisr.cpp:1373-1388), exact-rational bilinear ------------------------
struct ResizeParams {
    const void *in; size_t in_pitch; int in_w, in_h;
    void *out; size_t out_pitch; int W, H;
    const int *xmap, *xw, *ymap, *yw;
    int denx, deny;
};

template <typename PixT>
__global__ void __launch_bounds__(256) resize_kernel(const ResizeParams rp)
{
    PassParams p{};
    p.in = rp.in; p.in_pitch = rp.in_pitch; p.in_w = rp.in_w; p.in_h = rp.in_h; p.upscale = 1;
    p.xmap = rp.xmap; p.xw = rp.xw; p.ymap = rp.ymap; p.yw = rp.yw; p.denx = rp.denx; p.deny = rp.deny;
    const int X = blockIdx.x * 64 + (threadIdx.x & 63);
    const int Yb = blockIdx.y * 16 + (threadIdx.x >> 6) * 4;
    if (X >= rp.W) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int Y = Yb + k;
        if (Y >= rp.H) break;
        PixT *orow = reinterpret_cast<PixT *>(static_cast<char *>(rp.out) + (size_t)Y * rp.out_pitch);
        orow[X] = (PixT)(int)load_S<PixT>(p, Y, X);
    }
}

}  // namespace raisr
