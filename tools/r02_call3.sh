#!/bin/bash
# Round 2, GPU call 3: whole GPU suite, the BASELINE configurations through bench.py, ncu launch list with DRAM bytes.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gpu suite"
timeout 1100 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r02c_pytest.txt; grep -E "passed|failed|^FAILED|^ERROR|Error|assert " gpurun_out/r02c_pytest.txt | head -40
bench() { tag=$1; shift; echo "=== bench $tag: $*"; timeout 500 python bench.py "$@" > gpurun_out/r02c_bench_$tag.json 2> gpurun_out/r02c_bench_$tag.err; tail -1 gpurun_out/r02c_bench_$tag.json | cut -c1-400; tail -2 gpurun_out/r02c_bench_$tag.err; }
bench cfg2 --layers gpurun_out/r02c_layers_cfg2.tsv
bench cfg3 --config cfg3 --steps 10 --warmup 3 --cpu-images 2 --layers gpurun_out/r02c_layers_cfg3.tsv
bench cfg4 --config cfg4 --steps 10 --warmup 3 --cpu-images 3 --layers gpurun_out/r02c_layers_cfg4.tsv
bench cfg5 --config cfg5 --pool 2048 --warmup 3
echo "=== ncu launch list (time + DRAM bytes per launch), one 16-image step"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 420 \
    --csv --log-file gpurun_out/r02c_launches_step.csv python bench.py --steps 1 --warmup 1 --only-value > gpurun_out/r02c_ncu.log 2>&1
tail -3 gpurun_out/r02c_ncu.log; wc -l gpurun_out/r02c_launches_step.csv
