/*
 * x86_approx.c -- TEST INFRASTRUCTURE ONLY.  Executes the x86 approximation instructions the
 * reference's hash uses for "fast sqrt": vrcp14ps(vrsqrt14ps(x)) in the 16-wide AVX-512 hash
 * (Raisr_AVX512.cpp:200,221-222) and rcpps(rsqrtps(x)) in the 8-wide AVX2 hash it runs on the row
 * tails (Raisr_AVX256.cpp:419,441-442).  Used to pin the oracle against the compiled reference and
 * to generate/verify the lookup tables the CUDA engine ships for its "x86-exact" mode.
 */
#include "raisr_oracle.h"
#include <immintrin.h>

int oracle_have_x86_approx(void)
{
    return __builtin_cpu_supports("avx512f") ? 1 : 0;
}

__attribute__((target("avx512f"))) float oracle_x86_rcp14(float x)
{
    return _mm_cvtss_f32(_mm_rcp14_ss(_mm_setzero_ps(), _mm_set_ss(x)));
}
__attribute__((target("avx512f"))) float oracle_x86_rsqrt14(float x)
{
    return _mm_cvtss_f32(_mm_rsqrt14_ss(_mm_setzero_ps(), _mm_set_ss(x)));
}
float oracle_x86_rcpps(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }
float oracle_x86_rsqrtps(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }

/* bulk variants for table generation: n floats in, n floats out */
__attribute__((target("avx512f"))) void oracle_x86_rcp14_n(const float *x, float *y, long n)
{
    for (long i = 0; i < n; i++) y[i] = _mm_cvtss_f32(_mm_rcp14_ss(_mm_setzero_ps(), _mm_set_ss(x[i])));
}
__attribute__((target("avx512f"))) void oracle_x86_rsqrt14_n(const float *x, float *y, long n)
{
    for (long i = 0; i < n; i++) y[i] = _mm_cvtss_f32(_mm_rsqrt14_ss(_mm_setzero_ps(), _mm_set_ss(x[i])));
}
void oracle_x86_rcpps_n(const float *x, float *y, long n)
{
    for (long i = 0; i < n; i++) y[i] = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x[i])));
}
void oracle_x86_rsqrtps_n(const float *x, float *y, long n)
{
    for (long i = 0; i < n; i++) y[i] = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x[i])));
}
