// raisr_pipe_inst.cuh -- launch dispatch of the pipelined kernel for one (sample type, filter precision); included by raisr_pipe_*.cu
#pragma once
#include "raisr_launch.h"
#include "raisr_pipe_kernel.cuh"

namespace raisr {

template <typename PixT, int PT, int UA, int UB, int NV>
static int launch_one(const FrameLaunch &fl)
{
    auto kern = raisr_frame_pipe_kernel<PixT, PT, UA, UB, NV>;
    if (UB < 0) {
        kern<<<fl.grid, NTP, PIPE_SMEM_BYTES, fl.stream>>>(fl.a, fl.a);
        return (int)cudaGetLastError();
    }
    // chained passes: pass b's tiles spin on pass a's tile rows, so every CTA of the grid has to be resident -> cooperative launch
    void *args[2] = {const_cast<PassParams *>(&fl.a), const_cast<PassParams *>(&fl.b)};
    return (int)cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(kern), dim3(fl.grid), dim3(NTP), args, PIPE_SMEM_BYTES, fl.stream);
}

template <typename PixT, int PT, int UA, int UB, int NV>
static int prepare_one()
{
    return (int)cudaFuncSetAttribute(raisr_frame_pipe_kernel<PixT, PT, UA, UB, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM_BYTES);
}

// instantiated combinations: every single pass; chained pairs as the reference's two-pass modes produce them
//   mode 1: upscale in pass 1 (exact 2x with 4 pixel types, axis maps at 1.5x), none in pass 2;   mode 2: the other way round
#define RAISR_PIPE_COMBOS(X)                                                                                          \
    X(4, 0, -1) X(4, 1, -1) X(4, 2, -1) X(1, 0, -1) X(1, 1, -1) X(1, 2, -1)                                              \
    X(4, 1, 0) X(4, 0, 1) X(1, 2, 0) X(1, 0, 2)

template <typename PixT, int NV>
int launch_frame_pipe(const FrameLaunch &fl)
{
    const int pt = fl.a.ptypes, ua = fl.ups_a, ub = fl.two ? fl.ups_b : -1;
#define X(PT, UA, UB) if (pt == PT && ua == UA && ub == UB) return launch_one<PixT, PT, UA, UB, NV>(fl);
    RAISR_PIPE_COMBOS(X)
#undef X
    return (int)cudaErrorInvalidDeviceFunction;
}

template <typename PixT, int NV>
int prepare_frame_pipe()
{
    int rc = 0;
#define X(PT, UA, UB) if (!rc) rc = prepare_one<PixT, PT, UA, UB, NV>();
    RAISR_PIPE_COMBOS(X)
#undef X
    return rc;
}

}  // namespace raisr
