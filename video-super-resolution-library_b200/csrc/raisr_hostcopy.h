// raisr_hostcopy.h -- host-side row copies of the pageable-plane path (non-temporal stores where the CPU has them).
#pragma once
#include <cstddef>

namespace raisr {
// rows x row_bytes from src (row stride sstep) to dst (row stride dstep); the destination is written with streaming stores
void host_copy_rows(void *dst, size_t dstep, const void *src, size_t sstep, size_t row_bytes, int rows);
// 0 = memcpy, 1 = AVX2 streaming stores, 2 = AVX-512 streaming stores (decided once per process; RAISR_CUDA_NT_COPY=0 forces 0)
int host_copy_mode();
}  // namespace raisr
