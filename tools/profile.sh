#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list of one bench step + a full-set capture of conv kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/bench_under_ncu.log 2>&1
# full capture: a window of tcgen05 conv launches inside the first (warm-up) step's augmented pass
ncu --set full --clock-control none --import-source on -k regex:igemm_tc -s ${1:-100} -c ${2:-20} -o gpurun_out/prof_igemm $CMD > gpurun_out/prof.log 2>&1
ls -la gpurun_out
