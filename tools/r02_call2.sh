#!/bin/bash
# Round 2, GPU call 2: truncation pre-compensation (c sweep), stage errors, pool parity, whole GPU suite, bench.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 240 python tools/conv_accuracy.py 2.8e-8 3.2e-8 3.6e-8 > gpurun_out/r02b_convacc.txt 2>&1; grep -E "^conv|tensor core" gpurun_out/r02b_convacc.txt | head -60 )
( timeout 200 python tools/stage_error.py > gpurun_out/r02b_stage.txt 2>&1; tail -16 gpurun_out/r02b_stage.txt )
( timeout 300 python tools/pool_parity.py frcnn gpurun_out/pool_engine_frcnn.npz > gpurun_out/r02b_pool_frcnn.txt 2>&1; tail -10 gpurun_out/r02b_pool_frcnn.txt )
( timeout 400 python tools/pool_parity.py retina gpurun_out/pool_engine_retina.npz > gpurun_out/r02b_pool_retina.txt 2>&1; tail -10 gpurun_out/r02b_pool_retina.txt )
echo "=== gpu suite"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pool.py -x --maxfail=12 2>&1 | tail -60 > gpurun_out/r02b_pytest.txt; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r02b_pytest.txt | head -30
for cfg in "base:" "fuserpn:CALD_FUSE_RPN=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "=== bench $tag ($envs)"
  env $envs timeout 400 python bench.py --steps 10 --warmup 3 --cpu-images 1 --layers gpurun_out/r02b_layers_$tag.tsv > gpurun_out/r02b_bench_$tag.json 2> gpurun_out/r02b_bench_$tag.err
  tail -1 gpurun_out/r02b_bench_$tag.json | cut -c1-300
  tail -3 gpurun_out/r02b_bench_$tag.err
done
