#!/bin/bash
# Round 2, GPU call 17: ncu full-set capture of igemm_t_kernel with the four-half-stage ring (stem + layer1 3x3 launches)
mkdir -p gpurun_out /tmp/ncu
export PYTHONUNBUFFERED=1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:igemm_t_kernel -s 4 -c 4 -o /tmp/ncu/tform2 \
    python bench.py --steps 1 --warmup 1 --only-value --batch 16 > gpurun_out/r02q_ncu_tform.log 2>&1; tail -1 gpurun_out/r02q_ncu_tform.log
ncu -i /tmp/ncu/tform2.ncu-rep --page raw --csv > gpurun_out/r02q_tform_raw.csv 2>/dev/null
python tools/ncu_condense.py gpurun_out/r02q_tform_raw.csv > gpurun_out/r02q_tform_condensed.csv; cut -d, -f2,5,6,9,14,15 gpurun_out/r02q_tform_condensed.csv | cut -c1-200
