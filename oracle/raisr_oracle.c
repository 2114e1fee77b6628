/*
 * raisr_oracle.c -- CPU restatement of the reference's per-frame RAISR hot path (fp32 AVX512 flavour).
 *
 * TEST INFRASTRUCTURE ONLY (see raisr_oracle.h).  Plain C, IEEE arithmetic, no contraction
 * (built with -ffp-contract=off -fno-fast-math): every rounding below is one the reference performs.
 *
 * All file:line citations are relative to /root/reference/Library.
 */
#include "raisr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants (Raisr_globals.h:29, :208-264) ---------------------------------------------- */
static const float kPI = 3.141592653f;                      /* Raisr_globals.h:29 */

/* The 21 distinct literals of gGaussian2DOriginal (Raisr_globals.h:213-224); the 11x11 table is
 * symmetric in both axes and under transposition, so G[i][j] = kG[min(i',j')][max(i',j')] with
 * i' = min(i,10-i).  These are data and are used verbatim (never regenerated from exp()). */
static const double kG[6][6] = {
    {7.76554e-05, 0.000239195, 0.0005738, 0.001072, 0.00155975, 0.00176743},
    {0, 0.000736774, 0.00176743, 0.00330199, 0.00480437, 0.00544406},
    {0, 0, 0.00423984, 0.00792107, 0.0115251, 0.0130596},
    {0, 0, 0, 0.0147985, 0.0215317, 0.0243986},
    {0, 0, 0, 0, 0.0313284, 0.0354998},
    {0, 0, 0, 0, 0, 0.0402265}};

/* gGaussian2D{8,10,16}bit[i][lane]: lane 0 and lanes 12..15 are zero, lanes 1..11 carry
 * (float)(NF * literal) with NF a float expression (Raisr_globals.h:204-206). */
static void gaussian_table(int bits, float w[11][16])
{
    float M = bits == 8 ? 255.0f : bits == 10 ? 1023.0f : 65535.0f;
    float NF = 1.0f / (M * M * 2.0f * 2.0f);
    for (int i = 0; i < 11; i++) {
        int ii = i < 5 ? i : 10 - i;
        for (int l = 0; l < 16; l++) w[i][l] = 0.0f;
        for (int j = 0; j < 11; j++) {
            int jj = j < 5 ? j : 10 - j;
            int a = ii < jj ? ii : jj, b = ii < jj ? jj : ii;
            w[i][j + 1] = (float)((double)NF * kG[a][b]);
        }
    }
}

void oracle_gaussian_weights(int bits, float *w121)
{
    float w[11][16];
    gaussian_table(bits, w);
    for (int i = 0; i < 11; i++)
        for (int j = 0; j < 11; j++) w121[i * 11 + j] = w[i][j + 1];
}

/* ---- 16-lane horizontal sum, sumitup_ps_512 (Raisr_AVX512.cpp:37-44) ------------------------- */
static float sum16(const float *a)
{
    float t8[8], t4[4], t2[2];
    for (int j = 0; j < 8; j++) t8[j] = a[j] + a[j + 8];
    for (int j = 0; j < 4; j++) t4[j] = t8[j] + t8[j + 4];
    for (int j = 0; j < 2; j++) t2[j] = t4[j] + t4[j + 2];
    return t2[0] + t2[1];
}

/* ---- structure tensor of the pixel pair (r,cA),(r,cA+1): computeGTWG_Segment_AVX512_32f
 *      (Raisr_AVX512.cpp:69-131).  F = fp32 plane, stride W.  g[0..2] = pixel A, g[3..5] = pixel B. */
static void gtwg_pair(const float *F, int W, int r, int cA, const float wt[11][16], float g[6])
{
    float a0A[16] = {0}, a1A[16] = {0}, a2A[16] = {0};
    float a0B[16] = {0}, a1B[16] = {0}, a2B[16] = {0};
    const float *base = F + (size_t)(r - 6) * W + (cA - 6);
    for (int i = 0; i < 11; i++) {
        const float *ra = base + (size_t)i * W;         /* row r-6+i  */
        const float *rb = ra + W;                       /* row r-5+i  (the patch row) */
        const float *rc = rb + W;                       /* row r-4+i  */
        float gx[16], gy[16], va[16], vb[16], vc[16];
        /* The 16-float row loads of the reference run up to 2 floats past the end of a row for the right-most pairs
         * (and past the end of the buffer on the last hashed row).  Those lanes (14, 15) carry zero weight, so whatever
         * finite value is there contributes an exact +0; the restatement reads 0 instead of stray memory. */
        for (int j = 0; j < 16; j++) {
            const int inb = (cA - 6 + j) < W;
            va[j] = inb ? ra[j] : 0.0f; vb[j] = inb ? rb[j] : 0.0f; vc[j] = inb ? rc[j] : 0.0f;
        }
        for (int j = 0; j < 16; j++) {
            gx[j] = vc[j] - va[j];                                  /* GetGx: :54-57  */
            gy[j] = vb[(j + 1) & 15] - vb[(j + 15) & 15];           /* GetGy: :59-62 (lane rotates) */
        }
        for (int j = 0; j < 16; j++) {
            float wA = wt[i][j];
            float wB = wt[i][(j + 15) & 15];                        /* shiftR of the weight row, :111 */
            a0A[j] = fmaf(gx[j] * wA, gx[j], a0A[j]);               /* GetGTWG: fmadd(mul(a,w),b,acc) :64-67 */
            a1A[j] = fmaf(gx[j] * wA, gy[j], a1A[j]);
            a2A[j] = fmaf(gy[j] * wA, gy[j], a2A[j]);
            a0B[j] = fmaf(gx[j] * wB, gx[j], a0B[j]);
            a1B[j] = fmaf(gx[j] * wB, gy[j], a1B[j]);
            a2B[j] = fmaf(gy[j] * wB, gy[j], a2B[j]);
        }
    }
    g[0] = sum16(a0A); g[1] = sum16(a1A); g[2] = sum16(a2A);
    g[3] = sum16(a0B); g[4] = sum16(a1B); g[5] = sum16(a2B);
}

/* ---- hash (Raisr_AVX512.cpp:151-258 for 16-wide blocks, Raisr_AVX256.cpp:366-472 for 8-wide) - */
static float atan2_approx(float y, float x)     /* Raisr_AVX512.cpp:151-173 == Raisr_AVX256.cpp:366-391 */
{
    const float ONEQTR_PI = (float)(M_PI / 4.0);
    const float THRQTR_PI = (float)(3.0 * M_PI / 4.0);
    float ay = fabsf(y) + 1e-10f;
    int neg = x < 0.0f;
    float q = neg ? (x + ay) / (ay - x) : (x - ay) / (x + ay);
    float base = neg ? THRQTR_PI : ONEQTR_PI;
    float v = fmaf(fmaf(0.1963f * q, q, -0.9817f), q, base);
    return (y < 0.0f) ? -1.0f * v : v;
}

/* Bucket index from angle / strength / coherence.  wide16: thresholds <= value, NaN -> 0 (Raisr_AVX512.cpp:242-249);
 * 8-wide: 2 - [value <= Q0] - [value <= Q1], NaN -> 2 (Raisr_AVX256.cpp:457-464). */
static int quantise(float ang_scaled, float str, float coh, const oracle_pass_params *p, int wide16)
{
    float fa = floorf(ang_scaled);
    int ai;
    if (!(fa >= -2147483648.0f && fa < 2147483648.0f)) ai = (int)0x80000000; /* cvtps_epi32 of NaN/overflow */
    else ai = (int)fa;
    if (ai < 0) ai = 0;
    if (ai > 23) ai = 23;
    int si, ci;
    if (wide16) {
        si = (p->qstr[0] <= str) + (p->qstr[1] <= str);
        ci = (p->qcoh[0] <= coh) + (p->qcoh[1] <= coh);
    } else {
        si = 2 - ((str <= p->qstr[0]) + (str <= p->qstr[1]));
        ci = 2 - ((coh <= p->qcoh[0]) + (coh <= p->qcoh[1]));
    }
    return ai * 9 + si * 3 + ci;
}

/* ORACLE_SQRT_IEEE: the source semantics with IEEE sqrt and division.
 * wide16 != 0: GetHashValue_AVX512_32f_16Elements; == 0: GetHashValue_AVX256_32f_8Elements, which the
 * AVX512 build runs on the 8-wide tail blocks of every row (Raisr.cpp:1133-1134, 1247-1250). */
static int hash_bucket_ieee(const float g[3], const oracle_pass_params *p, int wide16)
{
    const float a = g[0], b = g[1], d = g[2];
    float T = a + d;
    float ad = a * d, bb = b * b;
    float D = ad - bb;
    float s = sqrtf((T * T) / 4.0f - D);
    float hT = T / 2.0f;
    float L1 = hT + s, L2 = hT - s;
    float x = (b != 0.0f) ? (L1 - d) : 1.0f;
    float ang = atan2_approx(b, x);
    ang = ang + ((ang < 0.0f) ? kPI : 0.0f);
    float s1 = sqrtf(L1), s2 = sqrtf(L2);
    float coh = (s1 - s2) / ((s1 + s2) + 0.00000000000000001f);
    const float qangle = 24.0f / kPI;                   /* gQAngle = gQuantizationAngle / PI, Raisr.cpp:1553 */
    return quantise(ang * qangle, L1, coh, p, wide16);
}

/* ORACLE_SQRT_X86: the same two functions AS COMPILED by g++ 13.3 with the reference's flag set
 * (-O3 -ffast-math -march=native with AVX-512; CMakeLists.txt:23-41), transcribed from the disassembly of
 * oracle/_ref/libraisr_ref.so (GetHashValue_AVX512_32f_16Elements at .text+0xa8e0; the 8-wide AVX2 variant
 * inlined into processSegment).  What -ffast-math did:
 *   - D and T*T/4 - D are contracted:  nD = fma(b,b,-(a*d));  z = fma(T*T, 0.25, nD)        (16-wide)
 *                                       z = (T*T)*NR(rcpps(4)) + nD  (two roundings)           (8-wide)
 *   - every division is reciprocal-approximation + one Newton step:  n/den = n * ((r+r) - r*(r*den)), r = rcp(den)
 *     (16-wide: vrcp14ps; 8-wide: vrcpps), EXCEPT the masked (x<0) atan2 branch of the 8-wide code, a true vdivps
 *   - T/2 is T*0.5 (exact) in the 16-wide code but T * NR(rcpps(2)) in the 8-wide code
 *   - gQAngle = 24 * (1/PI) (reciprocal constant), one ulp below 24/PI
 *   - sqrt is rcp(rsqrt(x)) as written in the source (Raisr_AVX512.cpp:200,221-222; Raisr_AVX256.cpp:419,441-442). */
static float nr_recip(float r, float den) { float e = r * (r * den); return (r + r) - e; }

static float atan_poly(float q, float base, float b)
{
    float v = fmaf(fmaf(q, 0.1963f * q, -0.9817f), q, base);
    return (b < 0.0f) ? -v : v;
}

static int hash_bucket_x86(const float g[3], const oracle_pass_params *p, int wide16)
{
    const float a = g[0], b = g[1], d = g[2];
    const float ONEQTR_PI = (float)(M_PI / 4.0), THRQTR_PI = (float)(3.0 * M_PI / 4.0);
    const float qangle = 24.0f * (1.0f / kPI);
    float T = a + d;
    float nD = fmaf(b, b, -(a * d));
    float ay = fabsf(b) + 1e-10f;
    float L1, L2, q, s1, s2, coh;
    if (wide16) {
        float z = fmaf(T * T, 0.25f, nD);
        float s = oracle_x86_rcp14(oracle_x86_rsqrt14(z));
        L1 = fmaf(T, 0.5f, s);
        L2 = fmaf(T, 0.5f, -s);
        float x = (b != 0.0f) ? (L1 - d) : 1.0f;
        float pl = x + ay, mn = x - ay, nd = ay - x;
        q = (x < 0.0f) ? pl * nr_recip(oracle_x86_rcp14(nd), nd) : mn * nr_recip(oracle_x86_rcp14(pl), pl);
        float ang = atan_poly(q, (x < 0.0f) ? THRQTR_PI : ONEQTR_PI, b);
        ang = ang + ((ang < 0.0f) ? kPI : 0.0f);
        s1 = oracle_x86_rcp14(oracle_x86_rsqrt14(L1));
        s2 = oracle_x86_rcp14(oracle_x86_rsqrt14(L2));
        float den = (s1 + s2) + 0.00000000000000001f;
        coh = (s1 - s2) * nr_recip(oracle_x86_rcp14(den), den);
        return quantise(ang * qangle, L1, coh, p, 1);
    } else {
        float quarter = nr_recip(oracle_x86_rcpps(4.0f), 4.0f);
        float half = nr_recip(oracle_x86_rcpps(2.0f), 2.0f);
        float z = (T * T) * quarter + nD;
        float s = oracle_x86_rcpps(oracle_x86_rsqrtps(z));
        float hT = T * half;
        L1 = s + hT;
        L2 = hT - s;
        float x = (b != 0.0f) ? (L1 - d) : 1.0f;
        float pl = x + ay, mn = x - ay, nd = ay - x;
        q = (x < 0.0f) ? pl / nd : mn * nr_recip(oracle_x86_rcpps(pl), pl);
        float ang = atan_poly(q, (x < 0.0f) ? THRQTR_PI : ONEQTR_PI, b);
        ang = ang + ((ang < 0.0f) ? kPI : 0.0f);
        s1 = oracle_x86_rcpps(oracle_x86_rsqrtps(L1));
        s2 = oracle_x86_rcpps(oracle_x86_rsqrtps(L2));
        float den = (s1 + s2) + 0.00000000000000001f;
        coh = (s1 - s2) * nr_recip(oracle_x86_rcpps(den), den);
        return quantise(ang * qangle, L1, coh, p, 0);
    }
}

static int hash_bucket(const float g[3], const oracle_pass_params *p, int wide16)
{
    return p->sqrt_mode == ORACLE_SQRT_X86 ? hash_bucket_x86(g, p, wide16) : hash_bucket_ieee(g, p, wide16);
}

/* ---- 121-tap filter, DotProdPatch_AVX512_32f (Raisr_AVX512.cpp:134-149) ---------------------- */
static float dot_patch(const float *F, int W, int r, int c, const float *f121)
{
    float acc[16];
    float pt[128], ft[128];
    for (int k = 0; k < 128; k++) {
        if (k < 121) {
            pt[k] = F[(size_t)(r - 5 + k / 11) * W + (c - 5 + k % 11)];
            ft[k] = f121[k];
        } else { pt[k] = 0.0f; ft[k] = 0.0f; }           /* rows are zero-padded to 128 (Raisr.cpp:299,329-331) */
    }
    for (int j = 0; j < 16; j++) acc[j] = pt[j] * ft[j];
    for (int m = 1; m < 8; m++)
        for (int j = 0; j < 16; j++) acc[j] = fmaf(pt[16 * m + j], ft[16 * m + j], acc[j]);
    return sum16(acc);
}

void oracle_hashed_cols(int W, int *c_end, int *tail_start)
{
    /* simulation of the column loop, Raisr.cpp:1065-1066 and 1246-1250, unrollSizePatchBased = 16 */
    int loopItr = 16, c = 6, tail = -1;
    while (c + loopItr <= W - 6) {
        if (loopItr == 8 && tail < 0) tail = c;
        if (loopItr > 8 && c + 2 * 16 > W - 6) loopItr = 8;
        c += loopItr;
    }
    if (tail < 0) tail = c;
    if (c_end) *c_end = c;
    if (tail_start) *tail_start = tail;
}

/* ---- census blend, CTCountOfBitsChangedSegment_AVX256_32f (Raisr_AVX256.cpp:68-166) ----------- */
static uint16_t blend_pixel(const float *L, const float *Hh, int W, int r, int c, int lo, int hi, int x86)
{
    const float lc = L[(size_t)r * W + c], hc = Hh[(size_t)r * W + c];
    int ham = 0;
    for (int i = -1; i <= 1; i++)
        for (int j = -1; j <= 1; j++) {
            if (i == 0 && j == 0) continue;
            int a = L[(size_t)(r + i) * W + c + j] < lc;
            int b = Hh[(size_t)(r + i) * W + c + j] < hc;
            ham += (a != b);
        }
    float w = (float)ham / 8.0f;
    float w2 = 1.0f - w;
    /* Source semantics (ORACLE_SQRT_IEEE): w*LR + (1-w)*HR, then + 0.5, each rounded.  As compiled with -ffast-math
     * (ORACLE_SQRT_X86; vector body and scalar tail alike, see the vfmadd132 pair before vrndscaleps in the reference
     * binary): fma(1-w, HR, fma(LR, w, 0.5)) -- w*LR + 0.5 is exact, so the whole sum is rounded once. */
    float v = x86 ? fmaf(w2, hc, fmaf(lc, w, 0.5f)) : (w * lc + w2 * hc) + 0.5f;
    float fv = floorf(v);
    int iv = (int)fv;
    if (iv > hi) iv = hi;
    if (iv < lo) iv = lo;
    return (uint16_t)iv;
}

/* ---- Randomness blend (blending = 1), Raisr.cpp:1203-1242 + CTRandomness_AVX512_32f (Raisr_AVX512.cpp:19-35) ----
 * Applied inside the hot loop to hashed pixels only, with the value of THIS evaluation (cur = filter result if in
 * range, else the upscale), so on the 16/8 overlap columns the 8-wide evaluation alone decides. */
static uint16_t randomness_pixel(const float *L, int W, int r, int c, float cur, int lo, int hi, int x86)
{
    const float lc = L[(size_t)r * W + c];
    int census = 0;
    for (int i = -1; i <= 1; i++)
        for (int j = -1; j <= 1; j++)
            if ((i || j) && L[(size_t)(r + i) * W + c + j] < lc) census++;
    float w = (float)census / 8.0f, w2 = 1.0f - w;
    /* source: weight*curPix + (1-weight)*LR, then += 0.5;  as compiled: fma(w, cur, w2*LR) (w2*LR is exact), then + 0.5 */
    float v = x86 ? fmaf(w, cur, w2 * lc) : (w * cur + w2 * lc);
    v += 0.5f;
    if (v < (float)lo) return (uint16_t)lo;
    if (v > (float)hi) return (uint16_t)hi;
    return (uint16_t)(int)v;
}

int oracle_pass(const uint16_t *S, int W, int H, const oracle_pass_params *p,
                int32_t *hash, float *gtwg, float *hr, uint16_t *out)
{
    if (!S || !out || !p || !p->filters || W < 1 || H < 1) return -1;
    if (p->blending != 1 && p->blending != 2) return -1;
    if (p->nptypes != 1 && p->nptypes != 4) return -1;
    const size_t N = (size_t)W * H;
    float *L = (float *)malloc(N * sizeof(float));
    float *Hh = (float *)malloc(N * sizeof(float));
    if (!L || !Hh) { free(L); free(Hh); return -1; }
    for (size_t i = 0; i < N; i++) { L[i] = (float)S[i]; Hh[i] = L[i]; }  /* convert + raisr32f := upscaled32f, Raisr.cpp:985-990,1029-1036 */
    if (hash) for (size_t i = 0; i < N; i++) hash[i] = -1;
    if (gtwg) memset(gtwg, 0, N * 3 * sizeof(float));

    float wt[11][16];
    gaussian_table(p->bits, wt);
    const float flo = (float)p->lo, fhi = (float)p->hi;
    /* Everything the blends below do not overwrite is the integer upscale itself, unclamped (row/edge memcpys,
     * Raisr.cpp:999-1028, 1252-1265).  With Randomness blending the reference never writes pixels
     * (H-7, [c_end, W-6)); this restatement defines them as the upscale too. */
    for (size_t i = 0; i < N; i++) out[i] = S[i];

    /* hot double loop, Raisr.cpp:1038-1066.  Blocks of 16, then blocks of 8 near the right edge; the first
     * 8-block starts 8 columns after the last 16-block, i.e. it REDOES that block's second half with the
     * 8-wide hash (c += loopItr uses the already reduced loopItr, Raisr.cpp:1247-1250). */
    for (int r = 6; r < H - 6; r++) {
        int loopItr = 16, c = 6;
        while (c + loopItr <= W - 6) {
            for (int pix = 0; pix < loopItr / 2; pix++) {
                float g[6];
                int cA = c + 2 * pix;
                gtwg_pair(L, W, r, cA, wt, g);
                for (int q = 0; q < 2; q++) {
                    int cc = cA + q;
                    int hv = hash_bucket(g + 3 * q, p, loopItr == 16);
                    int pt = p->nptypes == 4 ? (((r - 5) % 2) * 2 + ((cc - 5) % 2)) : 0;     /* Raisr.cpp:1068-1096 */
                    const float *f = p->filters + ((size_t)hv * p->nptypes + pt) * 121;
                    float cur = dot_patch(L, W, r, cc, f);
                    if (cur > flo && cur < fhi) Hh[(size_t)r * W + cc] = cur;                /* Raisr.cpp:1192-1196 */
                    else cur = L[(size_t)r * W + cc];
                    if (p->blending == 1) out[(size_t)r * W + cc] = randomness_pixel(L, W, r, cc, cur, p->lo, p->hi, p->sqrt_mode == ORACLE_SQRT_X86);
                    if (hash) hash[(size_t)r * W + cc] = hv;
                    if (gtwg) memcpy(gtwg + ((size_t)r * W + cc) * 3, g + 3 * q, 3 * sizeof(float));
                }
            }
            if (loopItr > 8 && c + 2 * 16 > W - 6) loopItr = 8;
            c += loopItr;
        }
    }
    if (hr) memcpy(hr, Hh, N * sizeof(float));

    if (p->blending == 2)
    for (int r = 1; r < H - 1; r++)
        for (int c = 1; c < W - 1; c++) out[(size_t)r * W + c] = blend_pixel(L, Hh, W, r, c, p->lo, p->hi, p->sqrt_mode == ORACLE_SQRT_X86);
    free(L); free(Hh);
    return 0;
}

/* ---- cheap upscale: restates oracle/ipp_standin/ipp.h (exact rational bilinear) --------------- */
static void axis_map(int d, int srcDim, int dstDim, int *i0, int *i1, long long *w1, long long *den)
{
    long long D = 2LL * dstDim;
    long long num = (2LL * d + 1) * srcDim - dstDim;
    long long q = num >= 0 ? num / D : -((-num + D - 1) / D);
    long long rr = num - q * D;
    int a = (int)q, b = (int)q + 1;
    if (a < 0) a = 0;
    if (a > srcDim - 1) a = srcDim - 1;
    if (b < 0) b = 0;
    if (b > srcDim - 1) b = srcDim - 1;
    *i0 = a; *i1 = b; *w1 = rr; *den = D;
}

void oracle_resize(const uint16_t *in, int inW, int inH, uint16_t *out, int outW, int outH)
{
    for (int y = 0; y < outH; y++) {
        int j0, j1; long long wy1, dy;
        axis_map(y, inH, outH, &j0, &j1, &wy1, &dy);
        for (int x = 0; x < outW; x++) {
            int i0, i1; long long wx1, dx;
            axis_map(x, inW, outW, &i0, &i1, &wx1, &dx);
            long long wy0 = dy - wy1, wx0 = dx - wx1, DD = dx * dy;
            long long s = wy0 * (wx0 * in[(size_t)j0 * inW + i0] + wx1 * in[(size_t)j0 * inW + i1]) +
                          wy1 * (wx0 * in[(size_t)j1 * inW + i0] + wx1 * in[(size_t)j1 * inW + i1]);
            out[(size_t)y * outW + x] = (uint16_t)((s + DD / 2) / DD);
        }
    }
}

int oracle_process_y(const uint16_t *in, int inW, int inH, uint16_t *out, int outW, int outH,
                     int passes, int mode, const oracle_pass_params *p1, const oracle_pass_params *p2,
                     int32_t *hash1, int32_t *hash2)
{
    /* exact ratios: (int)(outH / ratio) == inH */
    return oracle_process_y_ratio(in, inW, inH, out, outW, outH, (float)outH / (float)inH, passes, mode, p1, p2, hash1, hash2);
}

/* The reference builds the luma resize for {inW, (int)(segHeight / gRatio)} -> {outW, segHeight} (Raisr.cpp:1801-1803; one band:
 * segHeight = outH): when outH was truncated (odd input height at 1.5x: 135 -> 202, or vf_raisr's evenoutput) the vertical map is
 * built for FEWER source rows than the input plane has, and the last input row is never read. */
int oracle_process_y_ratio(const uint16_t *in, int inW, int inH, uint16_t *out, int outW, int outH, float ratio,
                           int passes, int mode, const oracle_pass_params *p1, const oracle_pass_params *p2,
                           int32_t *hash1, int32_t *hash2)
{
    int srcH = (int)((float)outH / ratio);
    if (srcH > inH) srcH = inH;
    if (srcH < 1) srcH = 1;
    if (passes != 1 && passes != 2) return -1;
    if (passes == 1) mode = 1;                          /* mode 2 is ignored with one pass, Raisr.cpp:1434-1435 */
    int rc = -1;
    uint16_t *up = NULL, *mid = NULL;
    if (passes == 1 || mode == 1) {
        up = (uint16_t *)malloc((size_t)outW * outH * sizeof(uint16_t));
        if (!up) return -1;
        oracle_resize(in, inW, srcH, up, outW, outH);
        if (passes == 1) { rc = oracle_pass(up, outW, outH, p1, hash1, NULL, NULL, out); goto done; }
        mid = (uint16_t *)malloc((size_t)outW * outH * sizeof(uint16_t));
        if (!mid) goto done;
        rc = oracle_pass(up, outW, outH, p1, hash1, NULL, NULL, mid);          /* quantised intermediate, Raisr.cpp:919-927 */
        if (rc == 0) rc = oracle_pass(mid, outW, outH, p2, hash2, NULL, NULL, out);
    } else {
        mid = (uint16_t *)malloc((size_t)inW * inH * sizeof(uint16_t));
        up = (uint16_t *)malloc((size_t)outW * outH * sizeof(uint16_t));
        if (!mid || !up) goto done;
        rc = oracle_pass(in, inW, inH, p1, hash1, NULL, NULL, mid);            /* pass 1 at input resolution */
        if (rc != 0) goto done;
        oracle_resize(mid, inW, srcH, up, outW, outH);                          /* pass 2 upscales, Raisr.cpp:945 */
        rc = oracle_pass(up, outW, outH, p2, hash2, NULL, NULL, out);
    }
done:
    free(up); free(mid);
    return rc;
}
