#!/bin/bash
# CTA-pair kernel: parity, micro-benchmark, whole-engine parity and bench A/B (run under gpurun)
mkdir -p gpurun_out
echo "== conv tests"; timeout 400 python -m pytest tests/test_gpu_conv.py -q -x 2>&1 | tail -5
echo "== micro"; timeout 300 python tools/pair_micro.py 2>&1 | grep -v "^$" | tail -30
echo "== full gpu suite, default"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
echo "== full gpu suite, CALD_CTA2=1"; CALD_CTA2=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
echo "== bench single"; CALD_CTA2=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_single.tsv > gpurun_out/bench_single.json 2>gpurun_out/bench_single.err; tail -1 gpurun_out/bench_single.json | cut -c1-200
echo "== bench pair"; CALD_CTA2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_pair.tsv > gpurun_out/bench_pair.json 2>gpurun_out/bench_pair.err; tail -1 gpurun_out/bench_pair.json | cut -c1-200
