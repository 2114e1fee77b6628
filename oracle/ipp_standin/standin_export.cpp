// standin_export.cpp -- TEST-ONLY: exports the IPP stand-in's resize (ipp.h, the code the compiled reference under
// oracle/_ref runs for its cheap upscale) so that bench.py's cpu_baseline context can time it on its own.
// Built by oracle/Makefile into _build/libipp_standin.so with the reference's flag set; never linked into the product.
#include "ipp.h"

extern "C" int standin_resize_8u(const unsigned char *src, int srcW, int srcH, int srcStep, unsigned char *dst, int dstW, int dstH, int dstStep)
{
    IppiResizeSpec_32f spec;
    ippiResizeLinearInit_8u({srcW, srcH}, {dstW, dstH}, &spec);
    int bufSize = 0;
    ippiResizeGetBufferSize_8u(&spec, {dstW, dstH}, 1, &bufSize);
    Ipp8u *buf = ippsMalloc_8u(bufSize);
    if (!buf) return ippStsNoMemErr;
    const int rc = ippiResizeLinear_8u_C1R(src, srcStep, dst, dstStep, {0, 0}, {dstW, dstH}, ippBorderRepl, 0, &spec, buf);
    ippsFree(buf);
    return rc;
}
