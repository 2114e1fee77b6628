/*
 * raisr/RaisrDefaults.h -- public value types of the RAISR library API.
 *
 * This is the drop-in boundary: names, field order, enum values and error codes are the ones the
 * reference publishes in Library/RaisrDefaults.h:10-57, because existing callers (ffmpeg/vf_raisr.c)
 * are compiled against them.  Nothing else is shared with the reference.
 */
#ifndef RAISR_B200_DEFAULTS_H
#define RAISR_B200_DEFAULTS_H

/* patch geometry of every shipped model (reference: RaisrDefaults.h:10-11) */
#define defaultPatchSize (11)
static const unsigned int defaultPatchAreaSize = defaultPatchSize * defaultPatchSize;

/* One image plane.  Caller-owned memory; `step` is the byte distance between rows (>= width * bytes
 * per sample).  `bitShift` is carried for ABI compatibility only (the reference reads it on its OpenCL
 * path alone, Raisr.cpp:1313-1348).  Reference: RaisrDefaults.h:13-20. */
typedef struct VideoDataType {
    unsigned char *pData;
    unsigned int width;
    unsigned int height;
    unsigned int step;
    unsigned int bitShift;
} VideoDataType;

/* Reference: RaisrDefaults.h:22-29 */
typedef enum RNLERRORTYPE {
    RNLErrorNone = 0,
    RNLErrorInsufficientResources = (int)0x80001000,
    RNLErrorUndefined = (int)0x80001001,
    RNLErrorBadParameter = (int)0x80001002,
    RNLErrorMax = (int)0x7FFFFFFF
} RNLERRORTYPE;

/* Reference: RaisrDefaults.h:31-35 */
typedef enum BlendingMode { Randomness = 1, CountOfBitsChanged = 2 } BlendingMode;

/* Instruction-set selector of the reference (RaisrDefaults.h:37-44).  The B200 engine has exactly one
 * compute path (sm_100a CUDA, fp32, AVX512-path numerics); every value is accepted and ignored. */
typedef enum ASMType { AVX2 = 1, AVX512 = 2, OpenCL = 3, OpenCLExternal = 4, AVX512_FP16 = 5 } ASMType;

/* Reference: RaisrDefaults.h:46-51 (kept so code that names it still compiles) */
typedef enum MachineVendorType { INTEL = 1, AMD = 2, VENDOR_UNSUPPORTED = 3 } MachineVendorType;

/* Reference: RaisrDefaults.h:53-57 */
typedef enum RangeType { VideoRange = 1, FullRange = 2 } RangeType;

#endif
