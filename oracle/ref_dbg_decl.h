/* TEST-ONLY: force-included (-include) when building oracle/_ref/libraisr_ref_dbg.so.
 * Planes the test harness points at [H*W] buffers to receive the reference's own
 * intermediate results (hash bucket, structure tensor) per pass. */
#pragma once
extern "C" {
inline int   *g_raisr_dbg_hash[2] = {nullptr, nullptr};   /* [pass][r*cols+c] */
inline float *g_raisr_dbg_gtwg[2] = {nullptr, nullptr};   /* [pass][(r*cols+c)*3+k] */
}
