#!/bin/bash
# Build libcald_b200.so for sm_100a (cross-compiles without a GPU).
#   ./build.sh            the product library (split-half operand planes, CALD_SPLIT_FP16=1)
#   ./build.sh ab         additionally libcald_b200_bf16.so (round 1's bfloat16 planes) for A/B measurements:
#                         CALD_LIB=libcald_b200_bf16.so selects it at load time (cald_b200/_lib.py)
set -e
cd "$(dirname "$0")"
SRCS="cald_b200/csrc/lib.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2 -shared -Iinclude"
nvcc $FLAGS -o cald_b200/libcald_b200.so $SRCS ${CALD_NVCC_EXTRA} &
P1=$!
if [ "$1" = "ab" ]; then
  nvcc $FLAGS -DCALD_SPLIT_FP16=0 -o cald_b200/libcald_b200_bf16.so $SRCS ${CALD_NVCC_EXTRA} &
  wait $!
fi
wait $P1
echo built cald_b200/libcald_b200.so
