#!/bin/bash
# Build libcald_b200.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=cald_b200/libcald_b200.so
SRCS="cald_b200/csrc/lib.cu"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
     -Xcompiler -fPIC,-O2 -shared -o $OUT $SRCS -Iinclude ${CALD_NVCC_EXTRA}
echo built $OUT
