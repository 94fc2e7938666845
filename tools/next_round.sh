#!/bin/bash
# First GPU call of the next round (run under gpurun, 1 GPU, ~6 min): the experiments that were built but not timed.
#  1. stage test of the epilogue-register shortcut (CALD_RESMMA_MAX_KB)
#  2. A/B of CALD_RESMMA_MAX_KB = unlimited vs 2 (GPU suite + bench, same box)
#  3. full-set ncu capture with SASS stall samples of layer3's expand convs (256 -> 1024 + shortcut): one-CTA launches
#     112..115 of `bench.py --steps 1 --warmup 1` (32 one-CTA conv launches per pass; the augmented pass starts at 96)
mkdir -p gpurun_out
CALD_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_detect.py -q -k "shortcut_in_epilogue or fused_rpn_head" 2>&1 | tail -3
bash tools/ab_measure.sh CALD_RESMMA_MAX_KB 1000000 2
#  2b. RPN 1x1 heads folded into the 3x3 RPN conv's epilogue (CALD_FUSE_RPN): bench off / on
for v in 0 1; do CALD_FUSE_RPN=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200; done
bash tools/epi_profile.sh 112 4
python tools/ncu_stalls.py gpurun_out/epi_source.csv 14 > gpurun_out/epi_stalls.txt; head -40 gpurun_out/epi_stalls.txt
