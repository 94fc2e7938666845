#!/bin/bash
# Full-set ncu capture (with source-level stall samples) of the first one-CTA conv launches of the augmented pass:
# layer1's short-K 1x1 convs, which are epilogue / HBM bound.  usage: tools/epi_profile.sh [skip] [count]
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workspace-gb 48"
ncu --set full --clock-control none --import-source on -k regex:igemm_tc_kernel -s ${1:-93} -c ${2:-8} -f -o /tmp/prof_epi $CMD > gpurun_out/epi_ncu.log 2>&1
ncu -i /tmp/prof_epi.ncu-rep --page raw --csv > /tmp/prof_epi_raw.csv 2>/dev/null
python tools/ncu_condense.py /tmp/prof_epi_raw.csv > gpurun_out/epi_capture.csv
ncu -i /tmp/prof_epi.ncu-rep --page source --csv > gpurun_out/epi_source.csv 2>/dev/null
ls -la gpurun_out | grep epi
