#!/bin/bash
# First GPU call of the next round (run under gpurun, 1 GPU, ~6 min): the experiments that were built but not timed.
#  1. stage test of the epilogue-register shortcut (CALD_RESMMA_MAX_KB)
#  2. A/B of CALD_RESMMA_MAX_KB = unlimited vs 2 (GPU suite + bench, same box)
#  3. full-set ncu capture with SASS stall samples of layer3's expand convs (256 -> 1024 + shortcut): one-CTA launches
#     112..115 of `bench.py --steps 1 --warmup 1` (32 one-CTA conv launches per pass; the augmented pass starts at 96)
mkdir -p gpurun_out
CALD_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_conv.py -q -k shortcut_in_epilogue 2>&1 | tail -3
bash tools/ab_measure.sh CALD_RESMMA_MAX_KB 1000000 2
bash tools/epi_profile.sh 112 4
python tools/ncu_stalls.py gpurun_out/epi_source.csv 14 > gpurun_out/epi_stalls.txt; head -40 gpurun_out/epi_stalls.txt
