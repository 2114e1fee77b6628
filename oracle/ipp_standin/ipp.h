/*
 * ipp.h -- TEST-ONLY stand-in for the ten Intel IPP entry points the reference
 * calls (Raisr.cpp:443-495, 950-957, 986-990, 1373-1388). Intel IPP is closed
 * source and is not in /root/reference, so the reference cannot be compiled
 * without something answering to this name.  This header exists ONLY so that
 * oracle/Makefile can build the untouched reference sources into
 * oracle/_ref/ (the checker / CPU baseline).  It is never included by the
 * product under video-super-resolution-library_b200/.
 *
 * Semantics the stand-in OWNS ("oracle policy"):
 *   - pixel-centre mapping   src = (dst + 0.5) * (srcDim / dstDim) - 0.5,
 *     with srcDim/dstDim taken from the spec exactly as IPP would;
 *   - replicate border (ippBorderRepl);
 *   - bilinear weights evaluated in EXACT rational arithmetic (integers over
 *     the denominator 2*dstDim per axis), result rounded half-up;
 *     at 2x this is (9a+3b+3c+d+8)>>4, at 1.5x it is (sum k*p + 18)/36.
 * This is the same statement of the stage as the reference's own OpenCL
 * "preprocess" kernel (Raisr_OpenCL_kernel.h:231-255: normalised (loc+0.5)*factor,
 * clamp-to-edge, linear, round).  A build against real IPP may differ by
 * +-1 LSB on some pixels (tie policy) -- parity at this stage is UNPINNED by
 * the reference and is pinned by this definition instead.
 */
#ifndef RAISR_ORACLE_IPP_STANDIN_H
#define RAISR_ORACLE_IPP_STANDIN_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char Ipp8u;
typedef unsigned short Ipp16u;
typedef int Ipp32s;
typedef unsigned int Ipp32u;
typedef float Ipp32f;
typedef int IppStatus;

enum { ippStsNoErr = 0, ippStsNoMemErr = -9 };
typedef enum { ippNearest = 1, ippLinear = 2, ippCubic = 6, ippLanczos = 16 } IppiInterpolationType;
typedef enum { ippBorderRepl = 1 } IppiBorderType;

typedef struct { int width; int height; } IppiSize;
typedef struct { int x; int y; } IppiPoint;

/* The "spec" only has to remember the two sizes the mapping is defined on. */
typedef struct IppiResizeSpec_32f {
    int srcW, srcH, dstW, dstH;
} IppiResizeSpec_32f;

static inline Ipp8u *ippsMalloc_8u(int len)
{
    void *p = NULL;
    if (len <= 0) len = 64;
    if (posix_memalign(&p, 64, (size_t)len) != 0) return NULL;
    return (Ipp8u *)p;
}
static inline void ippsFree(void *p) { free(p); }

static inline IppStatus ippiResizeGetSize_8u(IppiSize srcSize, IppiSize dstSize, IppiInterpolationType interp,
                                             Ipp32u antialiasing, int *pSpecSize, int *pInitBufSize)
{
    (void)srcSize; (void)dstSize; (void)interp; (void)antialiasing;
    *pSpecSize = (int)sizeof(IppiResizeSpec_32f);
    *pInitBufSize = 64;
    return ippStsNoErr;
}
static inline IppStatus ippiResizeLinearInit_8u(IppiSize srcSize, IppiSize dstSize, IppiResizeSpec_32f *pSpec)
{
    pSpec->srcW = srcSize.width; pSpec->srcH = srcSize.height;
    pSpec->dstW = dstSize.width; pSpec->dstH = dstSize.height;
    return ippStsNoErr;
}
static inline IppStatus ippiResizeLinearInit_16u(IppiSize srcSize, IppiSize dstSize, IppiResizeSpec_32f *pSpec)
{
    return ippiResizeLinearInit_8u(srcSize, dstSize, pSpec);
}
static inline IppStatus ippiResizeGetBufferSize_8u(const IppiResizeSpec_32f *pSpec, IppiSize dstSize,
                                                   Ipp32u numChannels, int *pBufSize)
{
    (void)pSpec; (void)numChannels;
    /* per-column (index, weight) tables for one call */
    *pBufSize = (int)((size_t)dstSize.width * 2 * sizeof(int) + 64);
    return ippStsNoErr;
}

/* One axis of the mapping: for destination index d returns the left/top source
 * index i0 (may be -1 .. srcDim-1 before clamping) and the weight numerator of
 * the SECOND tap over the denominator den = 2*dstDim:  pos = ((2d+1)*src - dst) / den. */
static inline void ipp_standin_axis(int d, int srcDim, int dstDim, int *i0, int *i1, int *w1, int *den)
{
    long long D = 2LL * dstDim;
    long long num = (2LL * d + 1) * srcDim - dstDim; /* may be negative */
    long long q = num >= 0 ? num / D : -((-num + D - 1) / D);
    long long r = num - q * D;
    int a = (int)q, b = (int)q + 1;
    if (a < 0) a = 0;
    if (a > srcDim - 1) a = srcDim - 1;
    if (b < 0) b = 0;
    if (b > srcDim - 1) b = srcDim - 1;
    *i0 = a; *i1 = b; *w1 = (int)r; *den = (int)D;
}

#define IPP_STANDIN_RESIZE(NAME, T)                                                                         \
    static inline IppStatus NAME(const T *pSrc, Ipp32s srcStep, T *pDst, Ipp32s dstStep, IppiPoint dstOffset, \
                                 IppiSize dstRoi, IppiBorderType border, const T *pBorderValue,              \
                                 const IppiResizeSpec_32f *pSpec, Ipp8u *pBuffer)                            \
    {                                                                                                        \
        (void)border; (void)pBorderValue;                                                                    \
        const int W = dstRoi.width, H = dstRoi.height;                                                       \
        int *xi = (int *)pBuffer; /* [W] packed i0 | (i1<<?) is overkill: store i0 and w1 */                 \
        int *xw = xi + W;                                                                                    \
        int denx = 1, deny = 1;                                                                              \
        const int exact2x = (pSpec->dstW == 2 * pSpec->srcW) && (pSpec->dstH == 2 * pSpec->srcH) &&          \
                            dstOffset.x == 0 && (dstOffset.y % 2) == 0;                                      \
        if (exact2x) {                                                                                       \
            /* fast path, identical arithmetic: weights {1/4,3/4}^2 -> (9a+3b+3c+d+8)>>4 */                  \
            const int sw = pSpec->srcW, sh = pSpec->srcH;                                                    \
            for (int y = 0; y < H; ++y) {                                                                    \
                int dy = y + dstOffset.y;                                                                    \
                int m = dy >> 1;                                                                             \
                int ya, yb, wa, wb; /* rows and integer weights /4 */                                        \
                if (dy & 1) { ya = m; yb = m + 1; wa = 3; wb = 1; } else { ya = m - 1; yb = m; wa = 1; wb = 3; } \
                if (ya < 0) ya = 0; if (yb > sh - 1) yb = sh - 1; if (ya > sh - 1) ya = sh - 1;              \
                const T *ra = (const T *)((const Ipp8u *)pSrc + (size_t)ya * srcStep);                       \
                const T *rb = (const T *)((const Ipp8u *)pSrc + (size_t)yb * srcStep);                       \
                T *out = (T *)((Ipp8u *)pDst + (size_t)y * dstStep);                                         \
                /* vertical pass into column sums v[x] = wa*ra[x]+wb*rb[x] (<= 4*65535) */                   \
                unsigned prev = wa * ra[0] + wb * rb[0]; /* replicate left */                                \
                unsigned cur = prev;                                                                         \
                for (int x = 0; x < sw; ++x) {                                                               \
                    unsigned nxt = (x + 1 < sw) ? (unsigned)(wa * ra[x + 1] + wb * rb[x + 1]) : cur;         \
                    if (2 * x < W) out[2 * x] = (T)((prev + 3 * cur + 8) >> 4);                              \
                    if (2 * x + 1 < W) out[2 * x + 1] = (T)((3 * cur + nxt + 8) >> 4);                       \
                    prev = cur; cur = nxt;                                                                   \
                }                                                                                            \
            }                                                                                                \
            return ippStsNoErr;                                                                              \
        }                                                                                                    \
        for (int x = 0; x < W; ++x) {                                                                        \
            int i0, i1, w1;                                                                                  \
            ipp_standin_axis(x + dstOffset.x, pSpec->srcW, pSpec->dstW, &i0, &i1, &w1, &denx);               \
            /* i1 is i0 or i0+1 after clamping; encode the step in the low bit */                            \
            xi[x] = (i0 << 1) | (i1 != i0);                                                                  \
            xw[x] = w1;                                                                                      \
        }                                                                                                    \
        for (int y = 0; y < H; ++y) {                                                                        \
            int j0, j1, wy1;                                                                                 \
            ipp_standin_axis(y + dstOffset.y, pSpec->srcH, pSpec->dstH, &j0, &j1, &wy1, &deny);              \
            const T *ra = (const T *)((const Ipp8u *)pSrc + (size_t)j0 * srcStep);                           \
            const T *rb = (const T *)((const Ipp8u *)pSrc + (size_t)j1 * srcStep);                           \
            T *out = (T *)((Ipp8u *)pDst + (size_t)y * dstStep);                                             \
            const long long wy0 = deny - wy1;                                                                \
            const long long DD = (long long)denx * deny;                                                     \
            for (int x = 0; x < W; ++x) {                                                                    \
                int i0 = xi[x] >> 1, i1 = i0 + (xi[x] & 1);                                                  \
                long long wx1 = xw[x], wx0 = denx - wx1;                                                     \
                long long s = wy0 * (wx0 * ra[i0] + wx1 * ra[i1]) + (long long)wy1 * (wx0 * rb[i0] + wx1 * rb[i1]); \
                out[x] = (T)((s + DD / 2) / DD);                                                             \
            }                                                                                                \
        }                                                                                                    \
        return ippStsNoErr;                                                                                  \
    }

IPP_STANDIN_RESIZE(ippiResizeLinear_8u_C1R, Ipp8u)
IPP_STANDIN_RESIZE(ippiResizeLinear_16u_C1R, Ipp16u)

static inline IppStatus ippiConvert_8u32f_C1R(const Ipp8u *pSrc, int srcStep, Ipp32f *pDst, int dstStep, IppiSize roi)
{
    for (int y = 0; y < roi.height; ++y) {
        const Ipp8u *s = pSrc + (size_t)y * srcStep;
        Ipp32f *d = (Ipp32f *)((Ipp8u *)pDst + (size_t)y * dstStep);
        for (int x = 0; x < roi.width; ++x) d[x] = (Ipp32f)s[x];
    }
    return ippStsNoErr;
}
static inline IppStatus ippiConvert_16u32f_C1R(const Ipp16u *pSrc, int srcStep, Ipp32f *pDst, int dstStep, IppiSize roi)
{
    for (int y = 0; y < roi.height; ++y) {
        const Ipp16u *s = (const Ipp16u *)((const Ipp8u *)pSrc + (size_t)y * srcStep);
        Ipp32f *d = (Ipp32f *)((Ipp8u *)pDst + (size_t)y * dstStep);
        for (int x = 0; x < roi.width; ++x) d[x] = (Ipp32f)s[x];
    }
    return ippStsNoErr;
}

#endif /* RAISR_ORACLE_IPP_STANDIN_H */
