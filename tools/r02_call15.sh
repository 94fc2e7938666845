#!/bin/bash
# Round 2, GPU call 15: the ncu launch list again with room for two whole 16-image steps; racecheck summary lines
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 \
    --csv --log-file gpurun_out/r02z_launches_step.csv python bench.py --steps 1 --warmup 1 --only-value --batch 16 > gpurun_out/r02z_ncu.log 2>&1
tail -2 gpurun_out/r02z_ncu.log; wc -l gpurun_out/r02z_launches_step.csv
timeout 400 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_conv.py -q -x -k "transposed and case0" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|igemm_t_kernel|Error:|Warning:" | cut -c1-300 | head -12 | tee gpurun_out/r02z_racecheck.txt
