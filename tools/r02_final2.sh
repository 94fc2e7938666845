#!/bin/bash
# Round 2, closing call: whole GPU suite on the final tree, then the BASELINE configurations through bench.py
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$|Warning|warnings.warn|cos = |sin = " > gpurun_out/r02y_pytest_full.txt; tail -2 gpurun_out/r02y_pytest_full.txt
bench() { tag=$1; shift; echo "=== bench $tag: $*"; ( time timeout 700 python bench.py "$@" > gpurun_out/r02y_bench_$tag.json 2> gpurun_out/r02y_bench_$tag.err ) 2>&1 | grep real; tail -1 gpurun_out/r02y_bench_$tag.json | cut -c1-200; tail -1 gpurun_out/r02y_bench_$tag.err | cut -c1-200; }
bench cfg2 --gpus 1 --steps 20 --warmup 5 --layers gpurun_out/r02y_layers_cfg2.tsv
bench cfg3 --config cfg3 --steps 10 --warmup 3 --cpu-images 2 --layers gpurun_out/r02y_layers_cfg3.tsv
bench cfg4 --config cfg4 --steps 10 --warmup 3 --cpu-images 3 --layers gpurun_out/r02y_layers_cfg4.tsv
