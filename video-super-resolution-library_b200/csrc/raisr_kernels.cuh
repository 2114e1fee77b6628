// raisr_kernels.cuh -- sm_100a device code of the RAISR luma pass and the chroma resize.
//
// One fused kernel per pass: a CTA owns a TW x th tile of the pass's output plane and runs, entirely out
// of shared memory,
//   A  cheap upscale (exact-rational bilinear) or plain load of the integer input -> S tile (+7 halo)
//   B  structure-tensor column chains, 8 rows at a time                            -> Q chunk
//   C  per pixel: 11-lane tree sums -> eigen-analysis -> bucket                    -> bucket tile (+1 halo)
//   D  per pixel type: TMA-bulk-load that type's 110 KB filter slice into shared memory, then the 121-tap
//      filter with 8 lanes per pixel (conflict-free 128-byte reads of the selected filter row)  -> HR tile
//   E  3x3 census blend, round, clamp, store
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
#include <string.h>
#include <type_traits>

namespace raisr {

struct PassParams {
    const void *in;          // integer input plane (u8/u16); LR-sized when upscale != 0
    size_t in_pitch;         // bytes
    int in_w, in_h;
    void *out;               // integer output plane, W x H
    size_t out_pitch;        // bytes
    int W, H;
    int row0, row1;          // output rows this launch produces: [row0, row1)
    int in_shift, out_shift; // samples carry their value in the HIGH bits (P010: 10 bits << 6): read as in >> in_shift, written as v << out_shift
    int upscale;             // 0: S = in,  1: S = resize(in)
    const int *xmap;         // [W] (i0 << 1 | step) : left source column, whether the right tap is i0+1
    const int *xw;           // [W] numerator of the right tap's weight over denx
    const int *ymap, *yw;    // same for rows, over deny
    int denx, deny;
    int up_src_h;            // source rows the vertical map was built for
    const float *filters;    // [ptypes][nbuckets][128]: per-type slices, rows lane-permuted for dot8()
    int ptypes;              // 4 or 1
    int nbuckets;            // 216
    int tile_h;              // output rows per CTA tile, <= TH_MAX (even)
    int vec_store;           // output base and pitch allow 4-pixel vector stores
    const unsigned *in_ready; // optional (pipelined kernel, split H2D): watermark of the input copies, (in_seq << 16) | rows that have landed
    unsigned in_seq;         // sequence number of this frame
    int in_split_row;        // input rows < in_split_row are valid by stream order; row r >= in_split_row once the watermark has passed r
    unsigned *err_flag;      // optional: set to 1 when an in-kernel flag wait times out (a copy the kernel waits for never landed)
    // chained two-pass launch (pipelined kernel): the first pass counts finished tiles per tile row, the second waits for the tile
    // rows of the first that cover the input rows it reads
    unsigned *rows_done;     // optional (first pass): [tile rows] finished tiles, zeroed by the host before the launch
    const unsigned *dep_done; // second pass: the first pass's rows_done
    int dep_gx, dep_ny;      //   tiles per tile row / tile rows of the first pass
    int dep_row0, dep_row1, dep_th;   // its output rows [dep_row0, dep_row1) and tile height
    unsigned *band_done;     // optional: per row band, the number of finished tiles (host-side copy pipeline)
    int band_tiles_y;        // tile rows per band
    void *out_tail;          // optional: output rows >= tail_row0 go here instead (the caller's pinned plane, written in place:
    size_t out_tail_pitch;   //           the rows of the last round of tiles need no copy after the kernel); not counted in band_done
    int tail_row0;
    float qstr0, qstr1, qcoh0, qcoh1;
    int lo, hi;              // colour range
    int c_end;               // hashed columns are [6, c_end)                  (Raisr.cpp:1065-1066)
    int tail_start;          // columns >= tail_start are hashed by the 8-wide variant (Raisr.cpp:1246-1250)
    int ov_end;              // columns in [tail_start, ov_end) are hashed by BOTH variants, 8-wide last
    int numerics;            // RAISR_NUMERICS_*
    float qangle;            // angle bins / PI (IEEE) or angle bins * (1/PI) (X86)
    int nangles;
    float quarter, half;     // X86 8-wide hash constants
    int *hash_out;           // optional [H][W] bucket plane (parity tests), -1 = not hashed
    int blending;            // 2 = CountOfBitsChanged
    // optional chroma planes resized by the pipelined kernel's producer warps once they have run out of luma tiles (the
    // filter warps are still busy with the last tile then): Raisr.cpp:1373-1388 without a launch of its own
    int chroma_n;            // 0, or the number of planes in chroma[]
    struct { const void *in; size_t in_pitch; void *out; size_t out_pitch; } chroma[2];
    int c_in_w, c_in_h, c_W, c_H;
    const int *c_xmap, *c_xw, *c_ymap, *c_yw;
    int c_denx, c_deny;
    int c_comps;                    // 1: planar chroma (chroma[0] = U, chroma[1] = V); 2: ONE semi-planar plane (NV12 / P010: U and V interleaved, chroma_n = 1)
    int c_shift;                    // chroma samples: value in the high bits (P010: 6), applied on load and store
    const unsigned *chroma_ready;   // optional: set to chroma_seq once the planes' H2D copies have landed
    unsigned chroma_seq;
    unsigned *chroma_done;          // optional: incremented once per CTA when its share of the chroma planes is written
    const uint2 *lut_rsqrt14, *lut_rcp14;      // x86 numerics: (c0,c1) runs of the 14-bit instructions, 128 entries each (may be null)
    const uint16_t *lut_rsqrtps, *lut_rcpps;   // x86 numerics: SSE approximation tables, 2048 entries each (may be null)
    // Gaussian weights of this engine's bit depth, folded: gw[i][m] = w[i][m] = w[i][10-m], m = 0..5 (Raisr_globals.h:204-264).
    // Part of the launch parameters (constant bank 0, read as uniform operands): per launch, hence per engine -- engines with
    // different bit depths on one device cannot disturb each other.
    alignas(8) float gw[11][6];
    // Pipelined kernel, exact-2x passes: CUtensorMap (opaque, 128 bytes) of the input plane, box = one tile's low-res window; stage A
    // of interior tiles fetches it with ONE cp.async.bulk.tensor (TMA) instead of ~2000 clamped scalar loads.  use_tmap = 0: scalar path.
    alignas(64) unsigned char in_tmap[128];
    int use_tmap;
};
constexpr int TMAP_BOX_H = 31;                                       // rows of the box: (PTH_MAX + 14) / 2 + 1 of the pipelined kernel
// box width: the window's first column is rounded down to a 16-byte boundary (TMA faults on a box whose first element is not 16-byte
// aligned in memory -- measured: "illegal instruction" for uint8 at column 100, fine at 96), so LRW (66) + up to 15 (uint8) / 7 (uint16)
// elements, rounded up to a multiple of 16 bytes
template <typename PixT> struct TmapBox { static constexpr int W = sizeof(PixT) == 1 ? 96 : 80, ALIGN = 16 / (int)sizeof(PixT); };

// ---- tile geometry -------------------------------------------------------------------------------
constexpr int NT = 512;          // threads per CTA (one CTA per SM: the filter slice alone is 110 KB)
constexpr int TW = 116;          // output tile width  (TW + 12 == 128 chain columns)
constexpr int TH_MAX = 62;       // output tile height is chosen per launch, <= TH_MAX (TH_MAX + 2 == 64 filtered rows)
constexpr int RB = 8;            // filtered rows per chunk
constexpr int QW = TW + 12;      // chain columns per row
constexpr int HW = TW + 2;       // filtered (HR) columns per row
constexpr int SW = TW + 14;      // S tile columns
constexpr int SP = 142;          // S tile pitch (floats): makes the 16 patch loads of dot8() bank-conflict free at column stride 1 and 2
constexpr int SH = TH_MAX + 14;  // S tile rows
constexpr int HP = HW + 2;       // HR tile pitch
constexpr int HH = TH_MAX + 2;   // HR tile rows
constexpr int GR = RB + 10;      // gradient rows per chunk
constexpr int NBUCKET_MAX = 216;
constexpr int SLICE_FLOATS = NBUCKET_MAX * 128;                 // one pixel type's filters, lane-permuted rows of 128 floats
constexpr int OVW = 8;                                          // width of the 16-wide/8-wide hash overlap (Raisr.cpp:1247-1250)
// shared memory carve-up (bytes).  The chunk buffers of stages B/C live inside the filter-slice buffer, which is
// only filled (by cp.async.bulk) once all chunks are done.
constexpr size_t OFF_S = 0;
constexpr size_t OFF_HR = OFF_S + sizeof(float) * SH * SP;
constexpr size_t OFF_F = (OFF_HR + sizeof(float) * HH * HP + 127) & ~(size_t)127;
constexpr size_t OFF_HASH = OFF_F + sizeof(float) * SLICE_FLOATS;
constexpr size_t OFF_HASH2 = OFF_HASH + (size_t)HH * HP;
constexpr size_t OFF_LUT = (OFF_HASH2 + (size_t)HH * OVW + 15) & ~(size_t)15;    // 256 x uint2: rsqrt14 runs, rcp14 runs
constexpr size_t OFF_MBAR = OFF_LUT + 2 * 128 * 4 * sizeof(unsigned);
constexpr size_t SMEM_BYTES = OFF_MBAR + 16;
static_assert(QW == 128 && HH % RB == 0 && SP >= SW && TW % 4 == 0 && HP % 4 == 0, "tile geometry");
constexpr int LRW = SW / 2 + 1, LRP = LRW + 1;   // low-res tile of the 2x fast path (lives in the slice buffer during stage A)
static_assert(sizeof(float) * (2u * GR * QW + (size_t)RB * 18 * QW) <= sizeof(float) * SLICE_FLOATS, "chunk buffers must fit in the slice buffer");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(OFF_F % 128 == 0, "slice buffer alignment");

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// ---- stage A: one sample of the cheap upscale (oracle/ipp_standin/ipp.h semantics) -----------------
template <typename PixT>
__device__ __forceinline__ float load_S(const PassParams &p, int Y, int X, bool upscale)
{
    if (!upscale) {
        const PixT *row = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)Y * p.in_pitch);
        return (float)((unsigned)row[X] >> p.in_shift);
    }
    const int xm = __ldg(p.xmap + X), ym = __ldg(p.ymap + Y);
    const int x0 = xm >> 1, x1 = x0 + (xm & 1), y0 = ym >> 1, y1 = y0 + (ym & 1);
    const int wx1 = __ldg(p.xw + X), wy1 = __ldg(p.yw + Y);
    const int wx0 = p.denx - wx1, wy0 = p.deny - wy1;
    const PixT *ra = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)y0 * p.in_pitch);
    const PixT *rb = reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)y1 * p.in_pitch);
    const unsigned a = (unsigned)ra[x0] >> p.in_shift, b = (unsigned)ra[x1] >> p.in_shift, c = (unsigned)rb[x0] >> p.in_shift, d = (unsigned)rb[x1] >> p.in_shift;
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

// the same sample for one component of a plane whose pixels are `comps` interleaved components (semi-planar chroma), value in the
// high bits (>> shift); axis maps and denominators from p
template <typename PixT>
__device__ __forceinline__ unsigned bilinear_sample(const PassParams &p, const void *in, size_t in_pitch, int Y, int X, int comps, int cc, int shift)
{
    const int xm = __ldg(p.xmap + X), ym = __ldg(p.ymap + Y);
    const int x0 = xm >> 1, x1 = x0 + (xm & 1), y0 = ym >> 1, y1 = y0 + (ym & 1);
    const int wx1 = __ldg(p.xw + X), wy1 = __ldg(p.yw + Y);
    const int wx0 = p.denx - wx1, wy0 = p.deny - wy1;
    const PixT *ra = reinterpret_cast<const PixT *>(static_cast<const char *>(in) + (size_t)y0 * in_pitch);
    const PixT *rb = reinterpret_cast<const PixT *>(static_cast<const char *>(in) + (size_t)y1 * in_pitch);
    const unsigned long long a = (unsigned)ra[x0 * comps + cc] >> shift, b = (unsigned)ra[x1 * comps + cc] >> shift;
    const unsigned long long c = (unsigned)rb[x0 * comps + cc] >> shift, d = (unsigned)rb[x1 * comps + cc] >> shift;
    const unsigned long long DD = (unsigned long long)p.denx * (unsigned long long)p.deny;
    const unsigned long long s = (unsigned long long)wy0 * ((unsigned long long)wx0 * a + (unsigned long long)wx1 * b) +
                                 (unsigned long long)wy1 * ((unsigned long long)wx0 * c + (unsigned long long)wx1 * d);
    return (unsigned)((s + DD / 2) / DD);
}

// ---- x86 approximation instructions via tables (numerics == X86) ------------------------------------
// vrsqrt14ps / vrcp14ps / rsqrtps / rcpps reproduced bit for bit (tables: tools/gen_x86_tables.py, verified there
// against the real instructions for every mantissa).  All four scale exactly with the exponent and depend on the top
// mantissa bits only (plus the exponent's parity for the square roots).  A table value v encodes the result for an
// argument in [1,2) (or [1,4)) as 0x3f000000 + (v << SHIFT).
//   14-bit forms: v = (c0 - c1 * low9) >> 9 with (c0, c1) selected by the top 7 (rcp) / parity + top 6 (rsqrt) bits;
//                 a zero mantissa returns an exact power of two.
//   SSE forms   : v = table[top 11 (rcp) / parity + top 10 (rsqrt) bits].
__host__ __device__ __forceinline__ unsigned f2u(float x)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    unsigned u; memcpy(&u, &x, 4); return u;
#endif
}
__host__ __device__ __forceinline__ float u2f(unsigned u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// (written select-style: the special cases are rare but must not cost divergent branches)
__host__ __device__ __forceinline__ float x86_rsqrt14(const uint2 *t, float x)
{
    const unsigned u = f2u(x);
    const unsigned e8 = (u >> 23) & 0xffu, m = u & 0x7fffffu;
    const unsigned par = (e8 & 1u) ^ 1u;                           // parity of the unbiased exponent (127 is odd)
    const int k = ((int)e8 - 127 - (int)par) >> 1;                 // x = [1,4) * 4^k
    const uint2 c = t[(par << 6) | (m >> 17)];
    const unsigned v = (c.x - c.y * ((m >> 8) & 0x1ffu)) >> 9;
    unsigned r = 0x3f000000u + (v << 7) - ((unsigned)k << 23);
    r = ((par | m) == 0u) ? ((unsigned)(127 - k) << 23) : r;       // exact power of 4 -> exact power of 2
    r = (e8 == 0u) ? 0x7f800000u : r;                              // +0 (and denormals, never produced here) -> +inf
    r = (u == 0x7f800000u) ? 0u : r;                               // +inf -> 0
    r = (u > 0x7f800000u) ? 0x7fc00000u : r;                       // NaN or negative -> NaN
    r = (u == 0x80000000u) ? 0xff800000u : r;                      // -0 -> -inf
    return u2f(r);
}

__host__ __device__ __forceinline__ float x86_rcp14(const uint2 *t, float x)
{
    const unsigned u = f2u(x);
    const unsigned sign = u & 0x80000000u, au = u & 0x7fffffffu;
    const unsigned e8 = au >> 23, m = au & 0x7fffffu;
    const uint2 c = t[m >> 16];
    const unsigned v = (c.x - c.y * ((m >> 7) & 0x1ffu)) >> 9;
    unsigned r = 0x7e800000u - (e8 << 23) + (v << 7);              // 0x3f000000 + (v << 7) - ((e8 - 127) << 23)
    r = (m == 0u) ? (0x7f000000u - (e8 << 23)) : r;                // exact power of 2
    r = (e8 == 0u) ? 0x7f800000u : r;                              // 0 -> inf
    r = (e8 == 255u) ? (m ? au : 0u) : r;                          // inf -> 0, NaN -> NaN
    return u2f(sign | r);
}

// Fast forms for the 16-wide hash (device hot path).  Same values as the general functions above on the inputs the
// hash can produce:
//   x86_sqrt14(z) = vrcp14ps(vrsqrt14ps(z)) for z = +0 (-> inf -> 0), z < 0 or NaN (-> NaN) and positive normal z;
//   x86_rcp14_pos(d) = vrcp14ps(d) for positive normal d.  The hash only divides by sums that are >= 1e-17 or NaN, and
//   whenever the divisor is NaN so is the dividend, hence the quotient is NaN whatever finite pattern this returns.
// The two 14-bit tables live in shared memory PACKED and REPLICATED: entry e of a table is the word ((c0 >> 6) << 10) | c1
// (c0 is a multiple of 64 and < 2^25, c1 < 1024: checked where the tables are uploaded) and is stored four times, at words
// 4e + r, r = 0..3; a lane reads replica r = lane & 3 (the table pointers of HashCtx already point at it).  One 32-bit load per
// lookup instead of a 64-bit one, and two lanes can only collide when they use the same replica AND their entries differ by a
// multiple of 8: the eight lookups per pixel took 11.6 M of the kernel's 127 M shared-memory wavefronts, 6.7 M of them bank
// conflicts (profiles/r2_pipe_kernel_summary.txt).
// Round 2, once the kernel had become bound by instruction issue and the bucket warps its critical path: the tables are stored
// UNPACKED again, entry e as the pair (c0, -c1), twice (replica r = lane & 1 at words 4e + 2r): one 64-bit load and one IMAD per
// lookup instead of a 32-bit load, two shifts / masks to unpack, a negation and the IMAD -- 4 instructions fewer per lookup, 32 per
// pixel -- at the price of the bank conflicts the packing had removed (the shared-memory pipe has room now).  Same 4 KB.
#ifndef RAISR_LUT_UNPACKED
#define RAISR_LUT_UNPACKED 1
#endif
constexpr int LUT_WORDS = 2 * 128 * 4;         // rsqrt14 then rcp14
#if RAISR_LUT_UNPACKED
constexpr int LUT_REPLICA_MASK = 1, LUT_REPLICA_WORDS = 2;
__device__ __forceinline__ unsigned lut14(const unsigned *t, unsigned entry, unsigned low9)
{
    const uint2 w = *reinterpret_cast<const uint2 *>(t + entry * 4u);
    return (w.x + w.y * low9) >> 9;            // c0 - c1 * low9 (mod 2^32; the difference is never negative)
}
__device__ __forceinline__ void lut14_fill(unsigned *dst, const uint2 *rsqrt14, const uint2 *rcp14, int tid, int nthreads)
{
    for (int i = tid; i < LUT_WORDS / 2; i += nthreads) {          // pair i = (table, entry, replica)
        const uint2 c = (i < LUT_WORDS / 4) ? rsqrt14[(i >> 1) & 127] : rcp14[(i >> 1) & 127];
        dst[2 * i] = c.x;
        dst[2 * i + 1] = 0u - c.y;
    }
}
#else
constexpr int LUT_REPLICA_MASK = 3, LUT_REPLICA_WORDS = 1;
__device__ __forceinline__ unsigned lut14(const unsigned *t, unsigned entry, unsigned low9)
{
    const unsigned w = t[entry * 4u];
    return (((w >> 10) << 6) - (w & 1023u) * low9) >> 9;
}
__device__ __forceinline__ void lut14_fill(unsigned *dst, const uint2 *rsqrt14, const uint2 *rcp14, int tid, int nthreads)
{
    for (int i = tid; i < LUT_WORDS; i += nthreads) {
        const uint2 c = (i < LUT_WORDS / 2) ? rsqrt14[(i >> 2) & 127] : rcp14[(i >> 2) & 127];
        dst[i] = ((c.x >> 6) << 10) | c.y;
    }
}
#endif

__device__ __forceinline__ float x86_rcp14_pos(const unsigned *t, float x)
{
    const unsigned u = __float_as_uint(x);
    const unsigned e8 = u >> 23, m = u & 0x7fffffu;
    const unsigned v = lut14(t, (m >> 16) & 127u, (m >> 7) & 0x1ffu);
    const unsigned r = (m == 0u) ? 0x7f000000u : (0x7e800000u + (v << 7));
    return __uint_as_float(r - (e8 << 23));
}

__device__ __forceinline__ float x86_sqrt14(const unsigned *trsq, const unsigned *trcp, float z)
{
    // y = vrsqrt14ps(z) for positive normal z
    const unsigned u = __float_as_uint(z);
    const unsigned e8 = (u >> 23) & 0xffu, m = u & 0x7fffffu;
    const unsigned par = (e8 & 1u) ^ 1u;
    const int k = ((int)e8 - 127 - (int)par) >> 1;
    const unsigned v = lut14(trsq, (par << 6) | (m >> 17), (m >> 8) & 0x1ffu);
    unsigned y = 0x3f000000u + (v << 7) - ((unsigned)k << 23);
    y = ((par | m) == 0u) ? ((unsigned)(127 - k) << 23) : y;
    float r = x86_rcp14_pos(trcp, __uint_as_float(y));          // y is a positive normal number
    r = (z == 0.0f) ? 0.0f : r;                                  // rsqrt14(0) = inf, rcp14(inf) = 0
    return (z >= 0.0f) ? r : __uint_as_float(0x7fc00000u);       // negative or NaN
}

__host__ __device__ __forceinline__ float x86_rsqrtps(const uint16_t *t, float x)
{
    const unsigned u = f2u(x);
    if (u == 0u) return u2f(0x7f800000u);
    if (u == 0x80000000u) return u2f(0xff800000u);
    if (u > 0x7f800000u) return u2f(0x7fc00000u);
    if (u == 0x7f800000u) return 0.0f;
    const int E = (int)(u >> 23) - 127;
    const unsigned m = u & 0x7fffffu;
    const int par = E & 1;
    const int k = (E - par) >> 1;
    return u2f(0x3f000000u + ((unsigned)t[(par << 10) | (m >> 13)] << 11) - ((unsigned)k << 23));
}

__host__ __device__ __forceinline__ float x86_rcpps(const uint16_t *t, float x)
{
    const unsigned u = f2u(x);
    const unsigned sign = u & 0x80000000u, au = u & 0x7fffffffu;
    if (au > 0x7f800000u) return x;
    if (au == 0u) return u2f(sign | 0x7f800000u);
    if (au == 0x7f800000u) return u2f(sign);
    const int E = (int)(au >> 23) - 127;
    return u2f(sign | (0x3f000000u + ((unsigned)t[(au & 0x7fffffu) >> 12] << 11) - ((unsigned)E << 23)));
}

struct HashCtx {
    float qstr0, qstr1, qcoh0, qcoh1;
    int numerics;
    float qangle;            // IEEE: angles / PI;  X86: angles * (1 / PI)   (what g++ -ffast-math emits for Raisr.cpp:1553)
    int nangles;
    float quarter, half;     // X86 8-wide hash: Newton-refined rcpps(4), rcpps(2)
    const unsigned *rsqrt14, *rcp14;       // shared memory: packed, replicated 14-bit instruction tables, this lane's replica
    const uint16_t *rsqrtps, *rcpps;       // global (row tails only)
    float fnangles;          // (float)nangles
};

// atan2 approximation, Raisr_AVX512.cpp:151-173 (== Raisr_AVX256.cpp:366-391), given the quotient q
__device__ __forceinline__ float atan_poly(float q, bool xneg, float b)
{
    const float ONEQTR_PI = 0.78539816339744830962f, THRQTR_PI = 2.35619449019234492885f;
    const float v = ffma(ffma(fmul(0.1963f, q), q, -0.9817f), q, xneg ? THRQTR_PI : ONEQTR_PI);
    return (b < 0.0f) ? -v : v;
}

__device__ __forceinline__ int quantise(const HashCtx &h, bool wide16, float ang, float str, float coh)
{
    ang = fadd(ang, (ang < 0.0f) ? 3.141592653f : 0.0f);                       // PI, Raisr_globals.h:29
    // floor, then clamp to [0, nangles - 1]; NaN -> 0 (cvt.rmi.s32.f32 saturates and turns NaN into 0): the reference's
    // min(max(cvt(floor(.)), 0), 23) with its NaN -> INT_MIN -> 0, without a branch
    const int ai = min(max(__float2int_rd(fmul(ang, h.qangle)), 0), h.nangles - 1);
    int idx = ai * 9;
    if (wide16) {            // thresholds <= value, NaN -> 0   (Raisr_AVX512.cpp:242-249)
        idx += (h.qstr0 <= str) ? 3 : 0;
        idx += (h.qstr1 <= str) ? 3 : 0;
        idx += (h.qcoh0 <= coh) ? 1 : 0;
        idx += (h.qcoh1 <= coh) ? 1 : 0;
    } else {                 // 2 - [value <= Q0] - [value <= Q1], NaN -> 2   (Raisr_AVX256.cpp:457-464)
        idx += 8;
        idx -= (str <= h.qstr0) ? 3 : 0;
        idx -= (str <= h.qstr1) ? 3 : 0;
        idx -= (coh <= h.qcoh0) ? 1 : 0;
        idx -= (coh <= h.qcoh1) ? 1 : 0;
    }
    return idx;
}

// reciprocal approximation + one Newton step, the form g++ -ffast-math gives every division of the hash
__device__ __forceinline__ float nr_recip(float r, float den) { return fsub(fadd(r, r), fmul(r, fmul(r, den))); }

// Bucket of one pixel from its structure tensor (a, b, d).
// WIDE16: GetHashValue_AVX512_32f_16Elements (Raisr_AVX512.cpp:175-258); else the 8-wide AVX2 variant the
// AVX-512 build runs on row tails (Raisr_AVX256.cpp:393-472; Raisr.cpp:1133-1134).
// numerics IEEE: the source semantics with sqrt.rn / div.rn (oracle: hash_bucket_ieee).
// numerics X86 : the functions as compiled by g++ 13.3 -O3 -ffast-math (oracle: hash_bucket_x86, transcribed from the
//                disassembly of the reference binary): contracted determinant, rcp(rsqrt()) roots, rcp+Newton divisions.
// NUMK: numerics known at compile time (1: X86, the IEEE code is not even instantiated -- no out-of-line sqrt / division calls in the
// bucket warps' loop, which lets ptxas keep the hash constants in uniform registers) or -1: decided by h.numerics at run time.
template <bool WIDE16, int NUMK = -1>
__device__ __forceinline__ int hash_bucket(const HashCtx &h, float a, float b, float d)
{
    const float T = fadd(a, d);
    const float ay = fadd(fabsf(b), 1e-10f);
    if (NUMK < 0 ? (h.numerics == 0) : (NUMK == 0)) {
        const float D = fsub(fmul(a, d), fmul(b, b));
        const float s = __fsqrt_rn(fsub(fmul(fmul(T, T), 0.25f), D));
        const float hT = fmul(T, 0.5f);
        const float L1 = fadd(hT, s), L2 = fsub(hT, s);
        const float x = (b != 0.0f) ? fsub(L1, d) : 1.0f;
        const bool neg = x < 0.0f;
        const float q = __fdiv_rn(neg ? fadd(x, ay) : fsub(x, ay), neg ? fsub(ay, x) : fadd(x, ay));
        const float s1 = __fsqrt_rn(L1), s2 = __fsqrt_rn(L2);
        const float coh = __fdiv_rn(fsub(s1, s2), fadd(fadd(s1, s2), 0.00000000000000001f));
        return quantise(h, WIDE16, atan_poly(q, neg, b), L1, coh);
    }
    const float nD = ffma(b, b, -fmul(a, d));
    float L1, L2, q, s1, s2, rden;
    bool neg;
    if (WIDE16) {
        const float z = ffma(fmul(T, T), 0.25f, nD);
        const float s = x86_sqrt14(h.rsqrt14, h.rcp14, z);
        L1 = ffma(T, 0.5f, s);
        L2 = ffma(T, 0.5f, -s);
        const float x = (b != 0.0f) ? fsub(L1, d) : 1.0f;
        neg = x < 0.0f;
        const float den = neg ? fsub(ay, x) : fadd(x, ay);
        q = fmul(neg ? fadd(x, ay) : fsub(x, ay), nr_recip(x86_rcp14_pos(h.rcp14, den), den));
        s1 = x86_sqrt14(h.rsqrt14, h.rcp14, L1);
        s2 = x86_sqrt14(h.rsqrt14, h.rcp14, L2);
        const float cden = fadd(fadd(s1, s2), 0.00000000000000001f);
        rden = nr_recip(x86_rcp14_pos(h.rcp14, cden), cden);
    } else {
        const float z = fadd(fmul(fmul(T, T), h.quarter), nD);
        const float s = x86_rcpps(h.rcpps, x86_rsqrtps(h.rsqrtps, z));
        const float hT = fmul(T, h.half);
        L1 = fadd(s, hT);
        L2 = fsub(hT, s);
        const float x = (b != 0.0f) ? fsub(L1, d) : 1.0f;
        neg = x < 0.0f;
        const float pl = fadd(x, ay);
        q = neg ? __fdiv_rn(pl, fsub(ay, x)) : fmul(fsub(x, ay), nr_recip(x86_rcpps(h.rcpps, pl), pl));
        s1 = x86_rcpps(h.rcpps, x86_rsqrtps(h.rsqrtps, L1));
        s2 = x86_rcpps(h.rcpps, x86_rsqrtps(h.rsqrtps, L2));
        const float cden = fadd(fadd(s1, s2), 0.00000000000000001f);
        rden = nr_recip(x86_rcpps(h.rcpps, cden), cden);
    }
    return quantise(h, WIDE16, atan_poly(q, neg, b), L1, fmul(fsub(s1, s2), rden));
}

// 11 lane values -> the reference's 16-lane tree (sumitup_ps_512, Raisr_AVX512.cpp:37-44).  Pixel A of a pair occupies
// lanes 1..11 and pixel B lanes 2..12 (Raisr_AVX512.cpp:107-114).  With P = k7+k3, Q = (k0+k8)+k4, R = (k1+k9)+k5,
// S = (k2+k10)+k6 the tree of A is t4 = [P,Q,R,S] -> (P+R)+(Q+S) and the tree of B is t4 = [S,P,Q,R] -> (S+Q)+(P+R):
// the same additions with commuted operands, i.e. the same fp32 result -- one function serves both.
__device__ __forceinline__ float tree_sum(const float *k)
{
    const float P = fadd(k[7], k[3]);
    const float Q = fadd(fadd(k[0], k[8]), k[4]);
    const float R = fadd(fadd(k[1], k[9]), k[5]);
    const float S = fadd(fadd(k[2], k[10]), k[6]);
    return fadd(fadd(P, R), fadd(Q, S));
}

// ---- mbarrier / bulk-copy primitives (sm_90+ PTX) ---------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 2-D tensor-map load (TMA): box at element coordinates (c0, c1) of the plane described by *tmap -> shared memory, completion on bar
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int c0, int c1, void *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The reference's 16 -> 1 lane tree (sumitup_ps_512, Raisr_AVX512.cpp:37-44) over the 8 lanes of a pixel, lane q holding
// chains 2q (a0) and 2q+1 (a1):  t8[j] = acc[j] + acc[j+8];  t4[j] = t8[j] + t8[j+4];  t2[j] = t4[j] + t4[j+2];  t2[0] + t2[1].
// Butterfly form: 4 shuffles instead of 6; every add has the same two operands as the reference's (fp add commutes).
__device__ __forceinline__ float tree8(float a0, float a1, int q)
{
    const bool hi = q >= 4;
    // lanes 0..3 end up with t8[2q] = a0[q] + a0[q+4], lanes 4..7 with t8[2(q-4)+1] = a1[q-4] + a1[q]
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? a0 : a1, 4, 8);
    float v = fadd(hi ? a1 : a0, recv);
    v = fadd(v, __shfl_xor_sync(0xffffffffu, v, 2, 8));       // t4[0] lanes 0,2 | t4[2] lanes 1,3 | t4[1] lanes 4,6 | t4[3] lanes 5,7
    v = fadd(v, __shfl_xor_sync(0xffffffffu, v, 1, 8));       // t2[0] lanes 0..3 | t2[1] lanes 4..7
    return fadd(v, __shfl_xor_sync(0xffffffffu, v, 4, 8));
}

// 121-tap filter in the reference's order -- 16 lane chains over 8 chunks of 16 taps, then the 16-lane tree
// (DotProdPatch_AVX512_32f, Raisr_AVX512.cpp:134-149) -- evaluated by 8 GPU lanes per pixel: lane q owns chains 2q
// and 2q+1.  frow = the pixel's filter row in the slice buffer, stored lane-permuted as [n][q][4] =
// {tap(16*2n + 2q), tap(16*2n + 2q+1), tap(16*(2n+1) + 2q), tap(16*(2n+1) + 2q+1)}, so the 8 lanes of a pixel read
// 128 contiguous bytes per n (conflict-free) and 4 pixels per warp proceed in lock step.
// sp = &S[r-5][c-5]; off[m][e] = offset of tap 16m + 2q + e inside the patch (0 with a zero coefficient for taps >= 121).
__device__ __forceinline__ float dot8(const float *sp, const float *frow, const int (&off)[8][2], int q)
{
    float a0 = 0.0f, a1 = 0.0f;
    const float4 *f4 = reinterpret_cast<const float4 *>(frow) + q;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const float4 f = f4[n * 8];
        const float p0 = sp[off[2 * n][0]], p1 = sp[off[2 * n][1]];
        const float p2 = sp[off[2 * n + 1][0]], p3 = sp[off[2 * n + 1][1]];
        if (n == 0) { a0 = fmul(p0, f.x); a1 = fmul(p1, f.y); }
        else { a0 = ffma(p0, f.x, a0); a1 = ffma(p1, f.y, a1); }
        a0 = ffma(p2, f.z, a0);
        a1 = ffma(p3, f.w, a1);
    }
    return tree8(a0, a1, q);          // same value in the pixel's 8 lanes
}

// ---- stage E: blend + store, one thread = 4 consecutive pixels of a row (a 3x6 window of S and of HR, one vector store)
// blending 2: CTCountOfBitsChangedSegment_AVX256_32f (Raisr_AVX256.cpp:68-166) over rows/cols [1, dim-1); the 1-px frame is
//             the integer upscale itself (Raisr.cpp:999-1028,1252-1265).
// blending 1: Randomness (Raisr.cpp:1203-1242, CTRandomness_AVX512_32f Raisr_AVX512.cpp:19-35) on hashed pixels only,
//             everything else is the integer upscale (border memcpys).
template <typename PixT, int BL, int NUMK = -1>
__device__ __forceinline__ void stage_blend_store_t(const PassParams &p, const float *sS, const float *sHR, const unsigned char *sHash,
                                                  int x0, int y0, int th, int t0, int nthreads)
{
    const bool ieee = NUMK < 0 ? (p.numerics == 0) : (NUMK == 0);
    const int W = p.W, H = p.H;
    for (int idx = t0; idx < th * (TW / 4); idx += nthreads) {
        const int ty = idx / (TW / 4), tx = (idx - ty * (TW / 4)) * 4;
        const int Y = y0 + ty, X = x0 + tx;
        if (Y >= p.row1 || Y >= H || X >= W) continue;
        float sw[3][6], hw[3][6];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const float *s = sS + (ty + 6 + dy) * SP + tx + 6;
            const float *hq = sHR + (ty + dy) * HP + tx;
#pragma unroll
            for (int dx = 0; dx < 6; ++dx) { sw[dy][dx] = s[dx]; hw[dy][dx] = hq[dx]; }
        }
        int iv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float lc = sw[1][e + 1], hcv = hw[1][e + 1];
            int r;
            if (BL == 2) {
                int ham = 0;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        if (dy == 1 && dx == 1) continue;
                        ham += ((sw[dy][e + dx] < lc) != (hw[dy][e + dx] < hcv));
                    }
                const float w = fmul((float)ham, 0.125f);
                // source semantics: (w*LR + (1-w)*HR) + 0.5; as compiled (-ffast-math): fma(1-w, HR, fma(LR, w, 0.5)), one rounding
                const float v = ieee ? fadd(fadd(fmul(w, lc), fmul(fsub(1.0f, w), hcv)), 0.5f)
                                                  : ffma(fsub(1.0f, w), hcv, ffma(lc, w, 0.5f));
                r = min(max((int)floorf(v), p.lo), p.hi);
                if (Y == 0 || Y == H - 1 || X + e == 0 || X + e == W - 1) r = (int)lc;   // 1-px frame
            } else {
                int census = 0;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        if (dy == 1 && dx == 1) continue;
                        census += (sw[dy][e + dx] < lc);
                    }
                const float w = fmul((float)census, 0.125f), w2 = fsub(1.0f, w);
                // source: w*cur + (1-w)*LR, += 0.5; as compiled: fma(w, cur, (1-w)*LR) then + 0.5   (hcv = this pixel's cur)
                const float v = fadd(ieee ? fadd(fmul(w, hcv), fmul(w2, lc)) : ffma(w, hcv, fmul(w2, lc)), 0.5f);
                r = (v < (float)p.lo) ? p.lo : ((v > (float)p.hi) ? p.hi : (int)v);
                if (sHash[(ty + 1) * HP + tx + 1 + e] == 255) r = (int)lc;               // not hashed: border copy of the upscale
            }
            iv[e] = r << p.out_shift;
        }
        const bool tail = p.out_tail != nullptr && Y >= p.tail_row0;
        PixT *orow = reinterpret_cast<PixT *>(static_cast<char *>(tail ? p.out_tail : p.out) + (size_t)Y * (tail ? p.out_tail_pitch : p.out_pitch)) + X;
        if (p.vec_store && X + 3 < W) {
            if (sizeof(PixT) == 1) *reinterpret_cast<uint32_t *>(orow) = (uint32_t)iv[0] | ((uint32_t)iv[1] << 8) | ((uint32_t)iv[2] << 16) | ((uint32_t)iv[3] << 24);
            else *reinterpret_cast<uint2 *>(orow) = make_uint2((uint32_t)iv[0] | ((uint32_t)iv[1] << 16), (uint32_t)iv[2] | ((uint32_t)iv[3] << 16));
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (X + e < W) orow[e] = (PixT)iv[e];
        }
    }
}

template <typename PixT, int NUMK = -1>
__device__ __forceinline__ void stage_blend_store(const PassParams &p, const float *sS, const float *sHR, const unsigned char *sHash,
                                                  int x0, int y0, int th, int t0, int nthreads)
{
    if (p.blending == 2) stage_blend_store_t<PixT, 2, NUMK>(p, sS, sHR, sHash, x0, y0, th, t0, nthreads);
    else stage_blend_store_t<PixT, 1, NUMK>(p, sS, sHR, sHash, x0, y0, th, t0, nthreads);
}

// UPS: 0 = the pass does not upscale, 1 = exact 2x (weights {1/4,3/4}^2 from a low-res tile in shared memory),
//      2 = any ratio through the per-axis tables (1.5x).   PT: pixel types (4 at 2x, else 1).
template <typename PixT, int PT, int UPS>
__global__ void __launch_bounds__(NT, 1) raisr_pass_kernel(const PassParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sS = reinterpret_cast<float *>(smem_raw + OFF_S);      // [SH][SP]  S rows  y0-7 .. , cols x0-7 .. x0+TW+6
    float *sHR = reinterpret_cast<float *>(smem_raw + OFF_HR);    // [HH][HP]  HR rows y0-1 .. , cols x0-1 .. x0+TW
    float *sF = reinterpret_cast<float *>(smem_raw + OFF_F);      // filter slice; before that: chunk buffers
    float *sGX = sF;                                              // [GR][QW]
    float *sGY = sGX + GR * QW;                                   // [GR][QW]
    float *sQ = sGY + GR * QW;                                    // [RB][18][QW]
    unsigned char *sHash = smem_raw + OFF_HASH;                   // [HH][HP] bucket, 255 = not hashed
    unsigned char *sHash2 = smem_raw + OFF_HASH2;                 // [HH][OVW] 16-wide bucket of overlap columns, 255 = same
    unsigned *sLut = reinterpret_cast<unsigned *>(smem_raw + OFF_LUT);  // packed + replicated rsqrt14, rcp14 (lut14_fill)
    void *mbar = smem_raw + OFF_MBAR;

    const int tid = threadIdx.x;
    if (p.numerics != 0) lut14_fill(sLut, p.lut_rsqrt14, p.lut_rcp14, tid, NT);
    const int th = p.tile_h;
    const int x0 = blockIdx.x * TW;
    const int y0 = p.row0 + blockIdx.y * th;
    const int W = p.W, H = p.H;
    const int hh = th + 2;                                        // filtered rows of this tile

    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- A: S tile ---------------------------------------------------------------------------------
    if (UPS == 1) {
        // low-res tile (replicate border = clamped coordinates), then every low-res sample's 3x3 neighbourhood yields the
        // 2x2 outputs (2j, 2j+1): even outputs weigh (j-1, j) by (1,3), odd outputs (j, j+1) by (3,1); all sums are exact
        // integers < 2^24, so fp32 evaluates (9a+3b+3c+d+8)>>4 exactly (oracle/ipp_standin/ipp.h).
        float *sL = sF;
        const int ly0 = (y0 - 8) >> 1, lx0 = (x0 - 8) >> 1;       // x0, y0 even
        const int lrh = (th + 14) / 2 + 1;
        for (int idx = tid; idx < lrh * LRW; idx += NT) {
            const int ly = idx / LRW, lx = idx - ly * LRW;
            const int yy = min(max(ly0 + ly, 0), p.up_src_h - 1), xx = min(max(lx0 + lx, 0), p.in_w - 1);
            sL[ly * LRP + lx] = (float)((unsigned)reinterpret_cast<const PixT *>(static_cast<const char *>(p.in) + (size_t)yy * p.in_pitch)[xx] >> p.in_shift);
        }
        __syncthreads();
        // block (bi,bj) centred on low-res (bi,bj) emits S rows 2bi-1, 2bi and cols 2bj-1, 2bj (tile-local)
        for (int idx = tid; idx < (lrh - 1) * (LRW - 1); idx += NT) {
            const int bi = idx / (LRW - 1) + 1, bj = idx - (bi - 1) * (LRW - 1) + 1;
            const float *l = sL + bi * LRP + bj;
            const float a0 = l[-LRP - 1], a1 = l[-LRP], a2 = l[-LRP + 1];
            const float b0 = l[-1], b1 = l[0], b2 = l[1];
            // even output row 2j: rows (j-1, j) x (1,3);  here row 2bi-2.. handled as: S row (2bi-1) is odd output of low-res row bi-1?
            // tile-local S row sy <-> Y = y0-7+sy; Y odd for even sy.  Output Y=2j+1 (odd) uses rows (j, j+1) x (3,1); Y=2j uses (j-1, j) x (1,3).
            // Low-res row index of l[0] is j = ly0 + bi.  S row for Y=2j   : sy = 2j - (y0-7) = 2bi - 1.  S row for Y=2j-1 (odd, rows (j-1,j) x (3,1)): sy = 2bi - 2.
            const float vo0 = ffma(3.0f, a0, b0), vo1 = ffma(3.0f, a1, b1), vo2 = ffma(3.0f, a2, b2);     // Y = 2j-1: 3*row(j-1) + row(j)
            const float ve0 = ffma(3.0f, b0, a0), ve1 = ffma(3.0f, b1, a1), ve2 = ffma(3.0f, b2, a2);     // Y = 2j  : row(j-1) + 3*row(j)
            // columns likewise: X = 2i-1 (odd): 3*col(i-1) + col(i);  X = 2i: col(i-1) + 3*col(i);  sx = 2bj-2, 2bj-1
            const int sy = 2 * bi - 2, sx = 2 * bj - 2;
            float *d = sS + sy * SP + sx;
            d[0] = floorf(fmul(fadd(ffma(3.0f, vo0, vo1), 8.0f), 0.0625f));
            d[1] = floorf(fmul(fadd(ffma(3.0f, vo1, vo0), 8.0f), 0.0625f));
            d[SP] = floorf(fmul(fadd(ffma(3.0f, ve0, ve1), 8.0f), 0.0625f));
            d[SP + 1] = floorf(fmul(fadd(ffma(3.0f, ve1, ve0), 8.0f), 0.0625f));
            (void)vo2; (void)ve2;
        }
    } else {
        for (int idx = tid; idx < (th + 14) * SW; idx += NT) {
            const int sy = idx / SW, sx = idx - sy * SW;
            const int Y = y0 - 7 + sy, X = x0 - 7 + sx;
            float v = 0.0f;
            if (Y >= 0 && Y < H && X >= 0 && X < W) v = load_S<PixT>(p, Y, X, UPS != 0);
            sS[sy * SP + sx] = v;
        }
    }
    __syncthreads();

    HashCtx hc{p.qstr0, p.qstr1, p.qcoh0, p.qcoh1, p.numerics, p.qangle, p.nangles, p.quarter, p.half, sLut + (tid & LUT_REPLICA_MASK) * LUT_REPLICA_WORDS, sLut + LUT_WORDS / 2 + (tid & LUT_REPLICA_MASK) * LUT_REPLICA_WORDS,
               p.lut_rsqrtps, p.lut_rcpps, (float)p.nangles};
    const float flo = (float)p.lo, fhi = (float)p.hi;

    for (int h0 = 0; h0 < hh; h0 += RB) {
        const int rfirst = y0 - 1 + h0;                       // frame row of the chunk's first filtered row
        const bool any_hashed = (rfirst + RB > 6) && (rfirst < H - 6) && (x0 - 1 + HW > 6) && (x0 - 1 < p.c_end);
        if (any_hashed) {
            // gradients for S rows h0+1 .. h0+RB+10, chain column q <-> S col q+1
            for (int idx = tid; idx < GR * QW; idx += NT) {
                const int g = idx / QW, q = idx - g * QW;
                const float *s = sS + (h0 + 1 + g) * SP + q + 1;
                sGX[idx] = fsub(s[SP], s[-SP]);               // GetGx: next row - previous row (Raisr_AVX512.cpp:54-57)
                sGY[idx] = fsub(s[1], s[-1]);                 // GetGy: right - left            (Raisr_AVX512.cpp:59-62)
            }
            __syncthreads();
            // ---- B: column chains ---------------------------------------------------------------------
            for (int it = tid; it < RB * QW; it += NT) {
                const int rl = it / QW, q = it - rl * QW;
                const int r = rfirst + rl;
                if (r < 6 || r >= H - 6 || h0 + rl >= hh) continue;
                float acc[6][3];
#pragma unroll
                for (int m = 0; m < 6; ++m) acc[m][0] = acc[m][1] = acc[m][2] = 0.0f;
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const float gx = sGX[(rl + i) * QW + q], gy = sGY[(rl + i) * QW + q];
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        const float w = p.gw[i][m];
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
        // ---- C: bucket ------------------------------------------------------------------------------
        for (int it = tid; it < RB * QW; it += NT) {
            const int rl = it / QW, j = it - rl * QW;
            const int h = h0 + rl;
            if (j >= HW || h >= hh) continue;
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
                    if (c < p.ov_end) {     // also hashed by the 16-wide block before (kept if the 8-wide result is out of range)
                        const int h16 = hash_bucket<true>(hc, g[0], g[1], g[2]);
                        if (h16 != hv) hv2 = h16;
                    }
                }
                if (p.hash_out && r >= p.row0 && r < p.row1 && j >= 1 && j <= TW) p.hash_out[(size_t)r * W + c] = hv;
            }
            sHash[h * HP + j] = (unsigned char)hv;
            if (c >= p.tail_start && c < p.tail_start + OVW) sHash2[h * OVW + (c - p.tail_start)] = (unsigned char)hv2;
            sHR[h * HP + j] = sS[(h + 6) * SP + j + 6];
        }
        __syncthreads();
    }

    // ---- D: 121-tap filter, one pixel type at a time ----------------------------------------------------------
    {
        constexpr int JS = (PT == 4) ? 2 : 1;                     // column/row stride between pixels of one type
        constexpr int U = 4;                                      // pixel groups per warp iteration (one group = 4 pixels x 8 lanes)
        constexpr int NCOLS = (HW + JS - 1) / JS;                 // upper bound of same-type columns in a tile row
        constexpr int NBLK = (NCOLS + 4 * U - 1) / (4 * U);       // warp iterations per row
        constexpr int ULAST = (NCOLS - 4 * U * (NBLK - 1) + 3) / 4; // groups in the last block of a row
        const int lane = tid & 31, warp = tid >> 5;
        const int g = lane >> 3, q = lane & 7;
        int off[8][2];
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 16 * m + 2 * q + e;
                off[m][e] = (k < 121) ? (k / 11) * SP + (k % 11) : 0;
            }
        const bool has_ov = (x0 - 1 + HW > p.tail_start) && (x0 - 1 < p.tail_start + OVW);
        const int slice_bytes = p.nbuckets * 128 * (int)sizeof(float);
        const float4 *sF4 = reinterpret_cast<const float4 *>(sF) + q;
        for (int t = 0; t < PT; ++t) {
            if (tid == 0) {
                fence_proxy_async();                              // generic-proxy accesses to the buffer are done (barrier above)
                mbar_expect_tx(mbar, (unsigned)slice_bytes);
                const char *src = reinterpret_cast<const char *>(p.filters) + (size_t)t * slice_bytes;
                const int piece = slice_bytes / 4;                // 4 bulk copies (each a multiple of 16 bytes)
                for (int i = 0; i < 4; ++i) bulk_g2s(reinterpret_cast<char *>(sF) + i * piece, src + i * piece, (unsigned)piece, mbar);
            }
            // pixels of type t inside the HR tile: frame parity (r-5)&1 == t>>1, (c-5)&1 == t&1   (Raisr.cpp:1068-1096)
            const int jfirst = (PT == 4) ? ((((x0 - 1 - 5) & 1) == (t & 1)) ? 0 : 1) : 0;
            const int hfirst = (PT == 4) ? ((((y0 - 1 - 5) & 1) == (t >> 1)) ? 0 : 1) : 0;
            const int nrows = (hh - hfirst + JS - 1) / JS;
            mbar_wait(mbar, (unsigned)(t & 1));
            // one warp iteration = UU groups of 4 pixels of one tile row; the last block of a row is shorter
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
                    const float cur = tree8(a0[u], a1[u], q);     // identical in all 8 lanes of the pixel
                    bool ok = (cur > flo) && (cur < fhi);         // strict range test, Raisr.cpp:1192-1196
                    float res = cur;
                    const int j = jb + 4 * JS * u;
                    if (has_ov) {                                 // tile holds columns hashed by both variants (rare path)
                        const int c = x0 - 1 + j;
                        const int hv2 = (j < HW && c >= p.tail_start && c < p.tail_start + OVW) ? sHash2[h * OVW + (c - p.tail_start)] : 255;
                        const bool need2 = (hv2 != 255) && !ok && p.blending == 2;   // Randomness blends this evaluation only
                        if (__any_sync(0xffffffffu, need2)) {
                            const float cur16 = dot8(sp + 4 * JS * u, sF + (hv2 == 255 ? 0 : hv2) * 128, off, q);
                            if (need2 && cur16 > flo && cur16 < fhi) { ok = true; res = cur16; }
                        }
                    }
                    if (q == 0 && hv[u] != 255 && ok) sHR[h * HP + j] = res;
                }
            };
            for (int it = warp; it < nrows * NBLK; it += NT / 32) {
                const int ri = it / NBLK, bi = it - ri * NBLK;
                const int h = hfirst + ri * JS;
                const int jb = jfirst + (bi * 4 * U + g) * JS;    // column of this lane's pixel in group u = 0; group u is 4*JS*u further
                if (bi < NBLK - 1) block(std::integral_constant<int, U>{}, h, jb);
                else block(std::integral_constant<int, ULAST>{}, h, jb);
            }
            __syncthreads();
        }
    }

    stage_blend_store<PixT>(p, sS, sHR, sHash, x0, y0, th, tid, NT);
    // row-band completion signal: the host's D2H stream waits (cuStreamWaitValue32) for all tiles of a band
    if (p.band_done) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(p.band_done + blockIdx.y / p.band_tiles_y, 1u);
        }
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
        orow[X] = (PixT)(int)load_S<PixT>(p, Y, X, true);
    }
}

}  // namespace raisr
