#!/bin/bash
# Round 2, GPU call 6: whole GPU suite on the final code, JPEG ingest throughput, smoke, default bench as the driver runs it.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gpu suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02f_pytest.txt; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r02f_pytest.txt | head -30
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== jpeg ingest"; timeout 600 python tools/jpeg_bench.py 256 2>&1 | tail -6 | tee gpurun_out/r02f_jpeg_bench.txt
echo "=== bench as the driver runs it"
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err ) 2>&1 | grep real
tail -1 gpurun_out/r02f_bench.json | cut -c1-400; tail -2 gpurun_out/r02f_bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err ) 2>&1 | grep real
tail -1 gpurun_out/r02f_bench_ref.json | cut -c1-300
