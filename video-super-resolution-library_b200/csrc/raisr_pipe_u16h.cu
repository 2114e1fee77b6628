// pipelined kernel, uint16_t samples, fp16 filter stage (opt-in numerics 3): see raisr_pipe_inst.cuh
#include "raisr_pipe_inst.cuh"
namespace raisr {
template int launch_frame_pipe<uint16_t, 1>(const FrameLaunch &);
template int prepare_frame_pipe<uint16_t, 1>();
}  // namespace raisr
