#!/bin/bash
# Round 2, GPU call 18: is the one-CTA MMA paced by the dependency on its accumulator columns?  igemm_t with one N = 256
# instruction per k-step against two N = 128 instructions on disjoint columns
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python tools/tform_micro.py 16 2>&1 | tail -9 | tee gpurun_out/r02r_tform_micro.txt
