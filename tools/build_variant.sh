#!/bin/bash
# Experiment builds of the hot kernel: recompiles raisr_pipe_u8x.cu with extra flags and links variants/libraisr_<name>.so from it plus
# the regular objects.   usage: tools/build_variant.sh <name> [nvcc flags...]     (run tools/kvariants.py on the GPU box afterwards)
set -e
name=$1; shift
PKG=$(cd "$(dirname "$0")/../video-super-resolution-library_b200" && pwd)
make -s -C "$PKG" >/dev/null
mkdir -p "$PKG/variants/$name"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -diag-suppress 549,177 -fmad=false -Xcompiler -fPIC \
    -Xcompiler -ffp-contract=off -ccbin /usr/bin/g++ -I"$PKG/../include" -I"$PKG/csrc" -Xptxas -v "$@" \
    -c "$PKG/csrc/raisr_pipe_u8x.cu" -o "$PKG/variants/$name/raisr_pipe_u8x.o" 2> "$PKG/variants/$name/ptxas.log" || (tail -5 "$PKG/variants/$name/ptxas.log"; exit 1)
objs=$(ls "$PKG"/build/*.o | grep -v raisr_pipe_u8x.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o "$PKG/variants/libraisr_$name.so" $objs "$PKG/variants/$name/raisr_pipe_u8x.o" \
    -lcudart_static -lpthread -ldl -lrt
grep -A1 "raisr_frame_pipe_kernelIhLi4ELi1ELin1ELi4" "$PKG/variants/$name/ptxas.log" | grep -E "spill" | head -1
echo "built variants/libraisr_$name.so"
