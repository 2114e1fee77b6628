// x86_tables.h -- tables that reproduce, bit for bit, the four x86 approximation instructions the compiled reference
// uses as "fast sqrt" and (through g++ -ffast-math) as divisions in its bucket hash:
//   vrsqrt14ps / vrcp14ps (16-wide AVX-512 hash)  -- architecturally defined: a 7-bit table with linear interpolation,
//                                                    stored as (c0, c1) pairs, value = (c0 - c1 * low9) >> 9
//   rsqrtps / rcpps       (8-wide AVX2 hash)      -- implementation specific (these are Intel's): direct tables
// Generated and exhaustively verified against the real instructions by tools/gen_x86_tables.py
// (-> x86_tables_data.cpp).  Without that file x86_tables() returns false and RAISR_NUMERICS_X86 is refused.
#pragma once
#include <cstddef>
#include <cstdint>

namespace raisr {
struct X86Tables {
    const uint32_t *rsqrt14;   // [2 parities][64 runs][2]
    const uint32_t *rcp14;     // [128 runs][2]
    const uint16_t *rsqrtps;   // [2 parities][1024]
    const uint16_t *rcpps;     // [2048]
};
constexpr int kRsqrt14Words = 2 * 64 * 2, kRcp14Words = 128 * 2, kRsqrtpsEntries = 2048, kRcppsEntries = 2048;
bool x86_tables(X86Tables *t);
}  // namespace raisr
