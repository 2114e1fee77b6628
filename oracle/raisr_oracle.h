/*
 * raisr_oracle.h -- CPU restatement of the reference's per-frame RAISR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under video-super-resolution-library_b200/ may include,
 * link or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do,
 * and only as the checker.
 *
 * What it restates (all file:line relative to /root/reference/Library):
 *   processSegment            Raisr.cpp:890-1289      stage order, borders, validity test, loop bounds
 *   computeGTWG_Segment_AVX512_32f  Raisr_AVX512.cpp:69-131   structure tensor, lane/tree summation order
 *   GetHashValue_AVX512_32f_16Elements  Raisr_AVX512.cpp:175-258  eigen-analysis -> bucket
 *   GetHashValue_AVX256_32f_8Elements   Raisr_AVX256.cpp:393-472  (used by the AVX512 build for the
 *                                                      8-wide tail blocks, Raisr.cpp:1133-1134,1247-1250)
 *   DotProdPatch_AVX512_32f   Raisr_AVX512.cpp:134-149 121-tap filter, 8x16-lane order
 *   CTCountOfBitsChangedSegment_AVX256_32f  Raisr_AVX256.cpp:68-166  census blend, round, clamp
 *   gGaussian2D{8,10,16}bit   Raisr_globals.h:208-264  weight literals
 *   cheap upscale             call sites Raisr.cpp:945-958; arithmetic owned by oracle/ipp_standin/ipp.h
 *
 * Parity status: PINNED against the compiled reference (oracle/_ref/libraisr_ref*.so built by
 * oracle/Makefile from the untouched sources) on seeded frames, see tests/test_oracle_vs_ref.py and
 * tests/golden/.  The cheap-upscale stage is third-party (Intel IPP, closed source, absent): that one
 * stage is "parity unpinned" by the reference and is DEFINED by the exact-rational bilinear here.
 */
#ifndef RAISR_ORACLE_H
#define RAISR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORACLE_SQRT_IEEE = 0,   /* the spec: IEEE sqrtf and '/'                                  */
    ORACLE_SQRT_X86  = 1    /* the hash AS COMPILED (g++ 13.3, the reference's -O3 -ffast-math flag set):
                               vrcp14(vrsqrt14(x)) / rcpps(rsqrtps(x)) square roots, reciprocal+Newton
                               divisions, contracted determinant -- executed with the real instructions
                               on the host CPU; bit-identical to oracle/_ref/libraisr_ref*.so */
};

typedef struct {
    int bits;               /* 8, 10 or 16 (selects the Gaussian normalisation NF_8/10/16)   */
    int lo, hi;             /* colour range (Raisr.cpp:1451-1468)                            */
    int nptypes;            /* 4 for ratio 2 (pixel types), 1 otherwise (Raisr.cpp:1477-1480) */
    const float *filters;   /* dense [216][nptypes][121]                                     */
    float qstr[2];
    float qcoh[2];
    int sqrt_mode;          /* ORACLE_SQRT_*                                                 */
    int blending;           /* 2 = CountOfBitsChanged (default), 1 = Randomness              */
} oracle_pass_params;

/* One full RAISR pass over an integer-valued plane S (W x H, dense stride W, values as uint16):
 *   hash  [H*W] int32, -1 where the pixel is not hashed          (nullable)
 *   gtwg  [H*W*3] float, structure tensor (a,b,d) per hashed pixel (nullable)
 *   hr    [H*W] float, filtered plane ("raisr32f")                (nullable)
 *   out   [H*W] uint16, final plane of this pass                  (required)
 * Returns 0, or -1 on bad arguments. */
int oracle_pass(const uint16_t *S, int W, int H, const oracle_pass_params *p,
                int32_t *hash, float *gtwg, float *hr, uint16_t *out);

/* Exact-rational bilinear resize with pixel-centre mapping and replicate border
 * (the definition in oracle/ipp_standin/ipp.h, restated). */
void oracle_resize(const uint16_t *in, int inW, int inH, uint16_t *out, int outW, int outH);

/* Whole luma pipeline: passes in {1,2}, mode in {1,2} (Raisr.cpp:896-975).
 * p1 = tables of pass 1, p2 = tables of pass 2 (ignored when passes == 1).
 * hash1/hash2 (nullable) receive the bucket planes of each pass (pass-1 plane is inW x inH in mode 2). */
int oracle_process_y(const uint16_t *in, int inW, int inH, uint16_t *out, int outW, int outH,
                     int passes, int mode, const oracle_pass_params *p1, const oracle_pass_params *p2,
                     int32_t *hash1, int32_t *hash2);
/* same with the reference's ratio made explicit: the resize reads (int)(outH / ratio) source rows (Raisr.cpp:1801-1803) */
int oracle_process_y_ratio(const uint16_t *in, int inW, int inH, uint16_t *out, int outW, int outH, float ratio,
                     int passes, int mode, const oracle_pass_params *p1, const oracle_pass_params *p2,
                     int32_t *hash1, int32_t *hash2);

/* gGaussian2D{8,10,16}bit as a dense 11x11 table (Raisr_globals.h:208-264) */
void oracle_gaussian_weights(int bits, float *w121);

/* hashed column range of a row: [6, *c_end), columns >= *tail_start use the 8-wide tail hash */
void oracle_hashed_cols(int W, int *c_end, int *tail_start);

/* x86 approximation instructions, executed for real (x86_approx.c). 0 if the host lacks AVX-512F. */
int   oracle_have_x86_approx(void);
float oracle_x86_rcp14(float x);
float oracle_x86_rsqrt14(float x);
float oracle_x86_rcpps(float x);
float oracle_x86_rsqrtps(float x);

#ifdef __cplusplus
}
#endif
#endif
