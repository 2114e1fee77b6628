// pipelined kernel, uint16_t samples, exact with the X86 numerics compiled in (NV = 4): see raisr_pipe_inst.cuh
#include "raisr_pipe_inst.cuh"
namespace raisr {
template int launch_frame_pipe<uint16_t, 4>(const FrameLaunch &);
template int prepare_frame_pipe<uint16_t, 4>();
}  // namespace raisr
