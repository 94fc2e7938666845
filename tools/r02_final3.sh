#!/bin/bash
# Round 2, last call: whole GPU suite and the driver-style cfg-2 bench on the final tree (elect.sync roles)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/r02x_pytest_full.txt
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --layers gpurun_out/r02x_layers_cfg2.tsv > gpurun_out/r02x_bench_cfg2.json 2> gpurun_out/r02x_bench_cfg2.err ) 2>&1 | grep real
tail -1 gpurun_out/r02x_bench_cfg2.json | cut -c1-200
