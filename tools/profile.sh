#!/bin/bash
# Run on the GPU box (under gpurun).  (1) ncu launch list of a bench run, (2) a full-set capture of a few conv
# launches around the FPN / RPN 3x3 convs at the P2 level of the augmented pass, exported as CSV on the box
# (the .ncu-rep with sources can exceed gpurun's 64 MiB return limit).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline --workspace-gb 12"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_b2.csv $CMD > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:igemm_tc -s ${1:-280} -c ${2:-5} -f -o /tmp/prof_igemm $CMD > gpurun_out/prof.log 2>&1
ncu -i /tmp/prof_igemm.ncu-rep --page raw --csv > gpurun_out/prof_igemm_raw.csv 2>/dev/null
ncu -i /tmp/prof_igemm.ncu-rep --page source --csv > gpurun_out/prof_igemm_source.csv 2>/dev/null
ls -la /tmp/prof_igemm.ncu-rep gpurun_out
SZ=$(stat -c %s /tmp/prof_igemm.ncu-rep)
if [ "$SZ" -lt 30000000 ]; then cp /tmp/prof_igemm.ncu-rep gpurun_out/; fi
