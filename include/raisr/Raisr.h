/*
 * raisr/Raisr.h -- C++ flavour of the API (std::string model path, default arguments), signatures as in
 * the reference's Library/Raisr.h:14-33.
 */
#ifndef RAISR_B200_RAISR_H
#define RAISR_B200_RAISR_H
#include <string>
#include <vector>
#include "RaisrDefaults.h"
#include "RaisrVersion.h"

RNLERRORTYPE RNLInit(std::string &modelPath, float ratio, unsigned int bitDepth = 8,
                     RangeType rangeType = VideoRange, unsigned int threadCount = 20, ASMType asmType = AVX512,
                     unsigned int passes = 1, unsigned int twoPassMode = 1);
RNLERRORTYPE RNLSetRes(VideoDataType *inY, VideoDataType *inCr, VideoDataType *inCb,
                       VideoDataType *outY, VideoDataType *outCr, VideoDataType *outCb);
RNLERRORTYPE RNLProcess(VideoDataType *inY, VideoDataType *inCr, VideoDataType *inCb,
                        VideoDataType *outY, VideoDataType *outCr, VideoDataType *outCb,
                        BlendingMode blendingMode = CountOfBitsChanged);
RNLERRORTYPE RNLSetOpenCLContext(void *context, void *device_id, int platformIndex, int deviceIndex);
RNLERRORTYPE RNLDeinit();
#endif
