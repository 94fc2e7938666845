#!/bin/bash
# Round 2, GPU call 8: pool ingest with the entropy walk on host threads (default) vs on the device -- parity tests of
# both placements and the files-vs-pixels scoring rate.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== jpeg tests (both walks)"
timeout 600 python -m pytest tests/test_gpu_jpeg.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02h_pytest_jpeg.txt
echo "=== jpeg ingest"
CALD_TRACE_JPEG=1 timeout 600 python tools/jpeg_bench.py 512 2>&1 | tail -30 | tee gpurun_out/r02h_jpeg_bench.txt
