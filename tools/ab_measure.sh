#!/bin/bash
# A/B of one engine switch (run under gpurun): conv stage tests, whole GPU suite, then bench with VAR=V0 and VAR=V1.
# usage: tools/ab_measure.sh CALD_FUSE_DS 0 1
VAR=${1:-CALD_FUSE_DS}; V0=${2:-0}; V1=${3:-1}
mkdir -p gpurun_out
echo "== conv tests"; timeout 400 python -m pytest tests/test_gpu_conv.py -q -x 2>&1 | tail -5
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
i=0
for v in $V0 $V1; do
  echo "== bench $VAR=$v"
  env $VAR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_ab$i.tsv > gpurun_out/bench_ab$i.json 2>gpurun_out/bench_ab$i.err
  tail -1 gpurun_out/bench_ab$i.json | cut -c1-200
  i=$((i+1))
done
