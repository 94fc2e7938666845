#!/bin/bash
# Round 2, GPU call 7: final verification -- whole suite, JPEG ingest after the Huffman-walk changes, driver-style bench
# with 32-image chunks, compute-sanitizer over the new kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gpu suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02g_pytest.txt; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r02g_pytest.txt | head -30
echo "=== jpeg ingest"; timeout 600 python tools/jpeg_bench.py 512 2>&1 | tail -4 | tee gpurun_out/r02g_jpeg_bench.txt
echo "=== bench as the driver runs it"
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --layers gpurun_out/r02g_layers.tsv > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err ) 2>&1 | grep real
tail -1 gpurun_out/r02g_bench.json | cut -c1-300; tail -2 gpurun_out/r02g_bench.err
bash tools/r02_sanitize.sh
