// pipelined kernel, uint16_t samples, fp32 filter stage (bit-exact): see raisr_pipe_inst.cuh
#include "raisr_pipe_inst.cuh"
namespace raisr {
template int launch_frame_pipe<uint16_t, false>(const FrameLaunch &);
template int prepare_frame_pipe<uint16_t, false>();
}  // namespace raisr
