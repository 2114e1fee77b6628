// pipelined kernel, uint8_t samples, fp32 filter stage (bit-exact): see raisr_pipe_inst.cuh
#include "raisr_pipe_inst.cuh"
namespace raisr {
template int launch_frame_pipe<uint8_t, false>(const FrameLaunch &);
template int prepare_frame_pipe<uint8_t, false>();
}  // namespace raisr
