// raisr_api.cpp -- the reference-facing API (RNL* C++ functions and RNLHandler_* C symbols) on top of the
// thin C ABI in include/raisr_cuda.h.  Like the reference (process-global state, Raisr_globals.h:140-203) there
// is one engine per process; the call protocol is Init -> SetRes -> Process* -> Deinit (vf_raisr.c:146,286-318,334).
#include <cstdlib>
#include <iostream>
#include <string>

#include "raisr/Raisr.h"
#include "raisr/RaisrHandler.h"
#include "raisr_cuda.h"

namespace {
raisr_cuda_engine *g_engine = nullptr;

int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}
}  // namespace

RNLERRORTYPE RNLInit(std::string &modelPath, float ratio, unsigned int bitDepth, RangeType rangeType,
                     unsigned int threadCount, ASMType asmType, unsigned int passes, unsigned int twoPassMode)
{
    (void)threadCount;   // CPU row-band count of the reference (Raisr.cpp:1637-1642): the grid replaces it
    (void)asmType;       // ISA selector of the reference (Raisr.cpp:1481-1528): one CUDA path here
    std::cout << "RAISR [version]:\tRAISR Native Lib v" << RAISR_VERSION_MAJOR << "." << RAISR_VERSION_MINOR
              << " (" << raisr_cuda_version() << ")" << std::endl;
    std::cout << "-------------------------------------------\n";
    if (g_engine) { raisr_cuda_destroy(g_engine); g_engine = nullptr; }
    raisr_cuda_config cfg{};
    cfg.model_path = modelPath.c_str();
    cfg.ratio = ratio;
    cfg.bit_depth = bitDepth;
    cfg.range_type = (int)rangeType;
    cfg.passes = passes;
    cfg.two_pass_mode = twoPassMode;
    cfg.device = env_int("RAISR_CUDA_DEVICE", -1);
    cfg.numerics = env_int("RAISR_CUDA_NUMERICS", RAISR_NUMERICS_X86_IF_AVAILABLE);
    cfg.keep_hash = env_int("RAISR_CUDA_KEEP_HASH", 0);
    const int rc = raisr_cuda_create(&cfg, &g_engine);
    if (rc == 0) std::cout << "ASM Type: CUDA sm_100a" << std::endl;
    return (RNLERRORTYPE)rc;
}

RNLERRORTYPE RNLSetRes(VideoDataType *inY, VideoDataType *inCr, VideoDataType *inCb, VideoDataType *outY,
                       VideoDataType *outCr, VideoDataType *outCb)
{
    if (!g_engine || !inY || !outY || !inCr || !outCr) return RNLErrorBadParameter;
    // The reference builds ONE chroma resize spec from inCr/outCr and applies it to both planes (Raisr.cpp:1805-1812,1373-1388), i.e.
    // it silently assumes Cb == Cr geometry -- true for every pixel format of vf_raisr.c:158-162.  Say so instead of assuming.
    if (inCb && (inCb->width != inCr->width || inCb->height != inCr->height)) {
        std::cout << "[RAISR ERROR] input Cb plane geometry differs from Cr" << std::endl;
        return RNLErrorBadParameter;
    }
    if (outCb && (outCb->width != outCr->width || outCb->height != outCr->height)) {
        std::cout << "[RAISR ERROR] output Cb plane geometry differs from Cr" << std::endl;
        return RNLErrorBadParameter;
    }
    return (RNLERRORTYPE)raisr_cuda_set_res(g_engine, inY->width, inY->height, outY->width, outY->height,
                                            inCr->width, inCr->height, outCr->width, outCr->height);
}

RNLERRORTYPE RNLProcess(VideoDataType *inY, VideoDataType *inCr, VideoDataType *inCb, VideoDataType *outY,
                        VideoDataType *outCr, VideoDataType *outCb, BlendingMode blendingMode)
{
    // null checks of the reference, Raisr.cpp:1297-1299 and 1358
    if (!inCr || !inCr->pData || !outCr || !outCr->pData || !inY || !inY->pData || !outY || !outY->pData)
        return RNLErrorBadParameter;
    if (!inCb || !inCb->pData || !outCb || !outCb->pData) return RNLErrorBadParameter;
    if (!g_engine) return RNLErrorBadParameter;
    return (RNLERRORTYPE)raisr_cuda_process_host(g_engine, inY->pData, inY->step, inCr->pData, inCr->step, inCb->pData,
                                                 inCb->step, outY->pData, outY->step, outCr->pData, outCr->step,
                                                 outCb->pData, outCb->step, (int)blendingMode);
}

RNLERRORTYPE RNLSetOpenCLContext(void *, void *, int, int) { return RNLErrorNone; }

RNLERRORTYPE RNLDeinit()
{
    raisr_cuda_destroy(g_engine);
    g_engine = nullptr;
    return RNLErrorNone;
}

extern "C" {

RNLERRORTYPE RNLHandler_Init(const char *modelPath, float ratio, unsigned int bitDepth, RangeType rangeType,
                             unsigned int threadCount, ASMType asmType, unsigned int passes, unsigned int twoPassMode)
{
    if (!modelPath) return RNLErrorBadParameter;
    std::string path(modelPath);
    return RNLInit(path, ratio, bitDepth, rangeType, threadCount, asmType, passes, twoPassMode);
}

RNLERRORTYPE RNLHandler_SetRes(VideoDataType *inY, VideoDataType *inCr, VideoDataType *inCb, VideoDataType *outY,
                               VideoDataType *outCr, VideoDataType *outCb)
{
    return RNLSetRes(inY, inCr, inCb, outY, outCr, outCb);
}

RNLERRORTYPE RNLHandler_Process(VideoDataType *inY, VideoDataType *inCr, VideoDataType *inCb, VideoDataType *outY,
                                VideoDataType *outCr, VideoDataType *outCb, BlendingMode blendingMode)
{
    return RNLProcess(inY, inCr, inCb, outY, outCr, outCb, blendingMode);
}

RNLERRORTYPE RNLHandler_SetOpenCLContext(void *context, void *device_id, int platformIndex, int deviceIndex)
{
    return RNLSetOpenCLContext(context, device_id, platformIndex, deviceIndex);
}

RNLERRORTYPE RNLHandler_Deinit(void) { return RNLDeinit(); }

}  // extern "C"
