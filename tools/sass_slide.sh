#!/bin/bash
# Rebuilds the engine and dumps the SASS of the hot instantiation (uint8, 4 pixel types, exact 2x, one pass, X86 numerics compiled in) to /tmp/k.sass
# (instruction text only), printing the spill line of ptxas.   usage: tools/sass_slide.sh
set -e
PKG=$(cd "$(dirname "$0")/../video-super-resolution-library_b200" && pwd)
make -s -C "$PKG" 2>&1 | tail
grep -A1 "raisr_frame_pipe_kernelIhLi4ELi1ELin1ELi4" "$PKG/build/raisr_pipe_u8x.ptxas.log" | tail -1
cuobjdump -sass -fun '_ZN5raisr23raisr_frame_pipe_kernelIhLi4ELi1ELin1ELi4EEEvNS_10PassParamsES1_' "$PKG/build/raisr_pipe_u8x.o" | grep -v "^\s*/\* 0x" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s*\/\*.*$//' > /tmp/k.sass
grep -n "SHFL.IDX" /tmp/k.sass | sed -n 2,3p
