// raisr_launch.h -- the seam between the host engine (raisr_engine.cu) and the translation units that hold the kernels
// (raisr_pipe_*.cu: one per sample type x filter precision, raisr_tile.cu), so that the 50-odd instantiations compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "raisr_kernels.cuh"

namespace raisr {

// One launch of the pipelined kernel: pass a, optionally chained with pass b (b reads the plane a writes).
struct FrameLaunch {
    PassParams a, b;
    bool two = false;          // b is valid: chained launch (cooperative: every CTA must be resident, b's tiles wait for a's)
    int ups_a = 0, ups_b = 0;  // upscale flavour per pass (0 none, 1 exact 2x, 2 axis maps)
    int grid = 0;              // CTAs (<= SMs)
    cudaStream_t stream = nullptr;
};

// cudaError_t as int; cudaErrorInvalidDeviceFunction = "no chained instantiation for this combination": launch the passes separately
// NV: numerics variant (0 exact, 1 fp16 filter stage, 2 separable fast hash)
template <typename PixT, int NV> int launch_frame_pipe(const FrameLaunch &fl);
// per-device opt-in to > 48 KB dynamic shared memory for every instantiation of that translation unit (call once per engine)
template <typename PixT, int NV> int prepare_frame_pipe();

template <typename PixT> int launch_pass_tile(const PassParams &p, int ups, dim3 grid, cudaStream_t s);
template <typename PixT> int prepare_pass_tile();

size_t pipe_smem_bytes();
int pipe_threads();
int pipe_tile_h_max();

}  // namespace raisr
