// pipelined kernel, uint8_t samples, separable fast hash (opt-in numerics 4): see raisr_pipe_inst.cuh
#include "raisr_pipe_inst.cuh"
namespace raisr {
template int launch_frame_pipe<uint8_t, 2>(const FrameLaunch &);
template int prepare_frame_pipe<uint8_t, 2>();
}  // namespace raisr
