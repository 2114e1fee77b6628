// pipelined kernel, uint16_t samples, exact: see raisr_pipe_inst.cuh
#include "raisr_pipe_inst.cuh"
namespace raisr {
template int launch_frame_pipe<uint16_t, 0>(const FrameLaunch &);
template int prepare_frame_pipe<uint16_t, 0>();
}  // namespace raisr
