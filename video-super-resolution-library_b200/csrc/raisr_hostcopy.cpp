// raisr_hostcopy.cpp -- the byte mover of the pageable-plane path (raisr_engine.cu: CopyPool).
//
// FFmpeg's software frames (what ffmpeg/vf_raisr.c hands RNLHandler_Process) are pageable: the engine DMAs from/to page-locked
// staging planes of its own and host threads move the bytes between those and the caller's planes.  That path is bound by the
// host's copy bandwidth (15.5 MB per 1080p->4K yuv420p frame), so the mover matters: every destination here is written once
// and not read again by this core (staging planes are read by the copy engine, the caller's output planes by the caller, later),
// which is the case for non-temporal stores -- no read-for-ownership of the destination line (2 instead of 3 bytes of memory
// traffic per byte copied) and no eviction of the caller's working set.  Full 64-byte lines with AVX-512, 32-byte stores with
// AVX2, plain memcpy() otherwise or when RAISR_CUDA_NT_COPY=0; picked once per process from the CPU's feature bits.
#include "raisr_hostcopy.h"

#include <immintrin.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace raisr {
namespace {

constexpr size_t kMinStream = 4096;   // below this the set-up is not worth it (and tiny planes stay in cache anyway)

__attribute__((target("avx512f"))) void stream_avx512(char *dp, const char *sp, size_t n)
{
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(dp) & 63)) & 63;
    memcpy(dp, sp, head);
    dp += head; sp += head; n -= head;
    const size_t lines = n / 64;
    size_t i = 0;
    for (; i + 4 <= lines; i += 4) {
        const __m512i a = _mm512_loadu_si512(sp), b = _mm512_loadu_si512(sp + 64), c = _mm512_loadu_si512(sp + 128), d = _mm512_loadu_si512(sp + 192);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dp), a);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dp + 64), b);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dp + 128), c);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dp + 192), d);
        sp += 256; dp += 256;
    }
    for (; i < lines; ++i) {
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dp), _mm512_loadu_si512(sp));
        sp += 64; dp += 64;
    }
    memcpy(dp, sp, n - lines * 64);
}

__attribute__((target("avx2"))) void stream_avx2(char *dp, const char *sp, size_t n)
{
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(dp) & 31)) & 31;
    memcpy(dp, sp, head);
    dp += head; sp += head; n -= head;
    const size_t blocks = n / 128;
    for (size_t i = 0; i < blocks; ++i) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(sp)), b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(sp + 32)),
                      c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(sp + 64)), d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(sp + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dp), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dp + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dp + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dp + 96), d);
        sp += 128; dp += 128;
    }
    memcpy(dp, sp, n - blocks * 128);
}

int pick_mode()
{
    if (const char *z = std::getenv("RAISR_CUDA_NT_COPY"))
        if (std::atoi(z) == 0) return 0;
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f")) return 2;
    if (__builtin_cpu_supports("avx2")) return 1;
    return 0;
}

inline void copy_span(int mode, char *dst, const char *src, size_t n)
{
    if (mode == 0 || n < kMinStream) memcpy(dst, src, n);
    else if (mode == 2) stream_avx512(dst, src, n);
    else stream_avx2(dst, src, n);
}

}  // namespace

int host_copy_mode()
{
    static const int mode = pick_mode();
    return mode;
}

void host_copy_rows(void *dst, size_t dstep, const void *src, size_t sstep, size_t row_bytes, int rows)
{
    const int mode = host_copy_mode();
    char *d = static_cast<char *>(dst);
    const char *s = static_cast<const char *>(src);
    if (dstep == row_bytes && sstep == row_bytes) copy_span(mode, d, s, row_bytes * static_cast<size_t>(rows));
    else
        for (int y = 0; y < rows; ++y) copy_span(mode, d + static_cast<size_t>(y) * dstep, s + static_cast<size_t>(y) * sstep, row_bytes);
    if (mode) _mm_sfence();   // the streamed lines are globally visible before the job reports completion
}

}  // namespace raisr
