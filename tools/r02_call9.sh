#!/bin/bash
# Round 2, GPU call 9: the new parity tests -- full-size pools against the unmodified reference, edge shapes, file ingest.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_score.py tests/test_gpu_jpeg.py -m gpu -q -s 2>&1 | grep -vE "^\s*$|Warning|warnings.warn" | tail -40 | tee gpurun_out/r02i_pytest_new.txt
