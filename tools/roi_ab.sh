#!/bin/bash
# RoIAlign kernel check (run under gpurun): detector stage tests, then the kernel's duration in a bench step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_detect.py tests/test_gpu_golden.py -q -x 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:roialign -c 4 --csv --log-file gpurun_out/roi_times.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workspace-gb 48 > /dev/null 2>&1
grep roialign gpurun_out/roi_times.csv | cut -d, -f5,15- | head
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
