#!/bin/bash
# Round 2, GPU call 10: diagnose the one full-size RetinaNet image above 1e-3 (which view, which stage)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/diag_fullsize.py retina 2 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/r02j_diag_retina2.txt
