// phase-sequential kernel (RAISR_CUDA_KERNEL=tile: the cross-check implementation) + facts about the pipelined kernel the host needs
#include "raisr_launch.h"
#include "raisr_pipe_kernel.cuh"

namespace raisr {

size_t pipe_smem_bytes() { return PIPE_SMEM_BYTES; }
int pipe_threads() { return NTP; }
int pipe_tile_h_max() { return PTH_MAX; }

template <typename PixT, int PT, int UPS>
static int tile_one(const PassParams &p, dim3 grid, cudaStream_t s)
{
    raisr_pass_kernel<PixT, PT, UPS><<<grid, NT, SMEM_BYTES, s>>>(p);
    return (int)cudaGetLastError();
}

template <typename PixT>
int launch_pass_tile(const PassParams &p, int ups, dim3 grid, cudaStream_t s)
{
    if (p.ptypes == 4) {
        if (ups == 0) return tile_one<PixT, 4, 0>(p, grid, s);
        if (ups == 1) return tile_one<PixT, 4, 1>(p, grid, s);
        return tile_one<PixT, 4, 2>(p, grid, s);
    }
    if (ups == 0) return tile_one<PixT, 1, 0>(p, grid, s);
    if (ups == 1) return tile_one<PixT, 1, 1>(p, grid, s);
    return tile_one<PixT, 1, 2>(p, grid, s);
}

template <typename PixT>
int prepare_pass_tile()
{
    int rc = 0;
#define X(PT, UPS) if (!rc) rc = (int)cudaFuncSetAttribute(raisr_pass_kernel<PixT, PT, UPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    X(4, 0) X(4, 1) X(4, 2) X(1, 0) X(1, 1) X(1, 2)
#undef X
    return rc;
}

template int launch_pass_tile<uint8_t>(const PassParams &, int, dim3, cudaStream_t);
template int launch_pass_tile<uint16_t>(const PassParams &, int, dim3, cudaStream_t);
template int prepare_pass_tile<uint8_t>();
template int prepare_pass_tile<uint16_t>();

}  // namespace raisr
