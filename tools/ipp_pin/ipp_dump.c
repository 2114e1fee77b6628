/*
 * ipp_dump.c -- dumps what REAL Intel IPP produces for the cheap-upscale stage, so that the stand-in that defines this stage
 * in this repository (oracle/ipp_standin/ipp.h; DESIGN.md section 2 "parity unpinned -- one stage") can be pinned.
 *
 * Needs oneAPI IPP (the reference pins 2021.12.1 / 2022.0).  NOT buildable in the development image (IPP is closed source and
 * absent); meant to be run once by anyone who has it:
 *
 *     source /opt/intel/oneapi/setvars.sh
 *     gcc -O2 tools/ipp_pin/ipp_dump.c -o ipp_dump -lippi -lipps -lippcore
 *     python tools/ipp_pin/make_ipp_golden.py ./ipp_dump          # writes tests/golden/ipp_*.npz
 *
 * Call sequence = the reference's (Raisr.cpp:435-500 ippInit, :945-990 resize calls): ippiResizeGetSize_8u,
 * ippiResizeLinearInit_8u / _16u, ippiResizeGetBufferSize_8u, ippiResizeLinear_8u_C1R / _16u_C1R with ippBorderRepl.
 *
 * usage: ipp_dump <bits: 8|16> <inW> <inH> <outW> <outH> <in.raw> <out.raw>       (raw little-endian planes, no padding)
 */
#include <stdio.h>
#include <stdlib.h>

#include <ipp.h>

int main(int argc, char **argv)
{
    if (argc != 8) {
        fprintf(stderr, "usage: %s bits inW inH outW outH in.raw out.raw\n", argv[0]);
        return 2;
    }
    const int bits = atoi(argv[1]), inW = atoi(argv[2]), inH = atoi(argv[3]), outW = atoi(argv[4]), outH = atoi(argv[5]);
    const int bps = bits == 8 ? 1 : 2;
    IppiSize srcSize = {inW, inH}, dstSize = {outW, outH};
    IppiPoint dstOffset = {0, 0};
    int specSize = 0, initSize = 0, bufSize = 0;
    if (ippiResizeGetSize_8u(srcSize, dstSize, ippLinear, 0, &specSize, &initSize) != ippStsNoErr) return 3;
    IppiResizeSpec_32f *spec = (IppiResizeSpec_32f *)ippsMalloc_8u(specSize);
    IppStatus st = bits == 8 ? ippiResizeLinearInit_8u(srcSize, dstSize, spec) : ippiResizeLinearInit_16u(srcSize, dstSize, spec);
    if (st != ippStsNoErr) return 4;
    if (ippiResizeGetBufferSize_8u(spec, dstSize, 1, &bufSize) != ippStsNoErr) return 5;
    Ipp8u *work = ippsMalloc_8u(bufSize);
    Ipp8u *in = (Ipp8u *)malloc((size_t)inW * inH * bps), *out = (Ipp8u *)malloc((size_t)outW * outH * bps);
    FILE *f = fopen(argv[6], "rb");
    if (!f || fread(in, bps, (size_t)inW * inH, f) != (size_t)inW * inH) return 6;
    fclose(f);
    if (bits == 8)
        st = ippiResizeLinear_8u_C1R(in, inW, out, outW, dstOffset, dstSize, ippBorderRepl, 0, spec, work);
    else
        st = ippiResizeLinear_16u_C1R((const Ipp16u *)in, inW * 2, (Ipp16u *)out, outW * 2, dstOffset, dstSize, ippBorderRepl, 0, spec, work);
    if (st != ippStsNoErr) return 7;
    f = fopen(argv[7], "wb");
    if (!f || fwrite(out, bps, (size_t)outW * outH, f) != (size_t)outW * outH) return 8;
    fclose(f);
    return 0;
}
