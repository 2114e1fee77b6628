/*
 * raisr/RaisrHandler.h -- the C entry points ffmpeg/vf_raisr.c binds (vf_raisr.c:146,286-318,334-337).
 * Same five symbols, argument order and return codes as the reference's Library/RaisrHandler.h:15-48;
 * behind them sits the B200 engine (include/raisr_cuda.h) instead of the AVX/IPP code.
 */
#ifndef RAISR_B200_HANDLER_H
#define RAISR_B200_HANDLER_H
#include "RaisrDefaults.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Loads the model folder, selects the device, builds the engine.  threadCount and asmType are CPU
 * notions and are ignored.  Replaces RaisrHandler.h:15-23 / Raisr.cpp:1409-1679. */
RNLERRORTYPE RNLHandler_Init(const char *modelPath, float ratio, unsigned int bitDepth, RangeType rangeType,
                             unsigned int threadCount, ASMType asmType, unsigned int passes,
                             unsigned int twoPassMode);

/* Fixes the frame geometry and allocates device planes.  Called once, before the first Process.
 * Replaces RaisrHandler.h:25-31 / Raisr.cpp:1681-1829. */
RNLERRORTYPE RNLHandler_SetRes(VideoDataType *inY, VideoDataType *inCr, VideoDataType *inCb,
                               VideoDataType *outY, VideoDataType *outCr, VideoDataType *outCb);

/* One frame, blocking: host planes in, host planes out.  Replaces RaisrHandler.h:33-40 / Raisr.cpp:1294-1397. */
RNLERRORTYPE RNLHandler_Process(VideoDataType *inY, VideoDataType *inCr, VideoDataType *inCb,
                                VideoDataType *outY, VideoDataType *outCr, VideoDataType *outCb,
                                BlendingMode blendingMode);

/* Kept for link compatibility (RaisrHandler.h:42-46); there is no OpenCL here, returns RNLErrorNone. */
RNLERRORTYPE RNLHandler_SetOpenCLContext(void *context, void *device_id, int platformIndex, int deviceIndex);

/* Frees every device and host resource.  Replaces RaisrHandler.h:48 / Raisr.cpp:1842-1909. */
RNLERRORTYPE RNLHandler_Deinit(void);

#ifdef __cplusplus
}
#endif
#endif
