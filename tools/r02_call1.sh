#!/bin/bash
# Round 2, GPU call 1 (one B200): accuracy of the two operand formats, pool-level parity, the GPU suite, A/B benches.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
run() { echo "=== $*"; "$@"; }
export PYTHONUNBUFFERED=1
( timeout 300 python tools/conv_accuracy.py > gpurun_out/r02_convacc_fp16.txt 2>&1; tail -3 gpurun_out/r02_convacc_fp16.txt )
( CALD_LIB=libcald_b200_bf16.so timeout 300 python tools/conv_accuracy.py > gpurun_out/r02_convacc_bf16.txt 2>&1; tail -3 gpurun_out/r02_convacc_bf16.txt )
( timeout 300 python tools/stage_error.py > gpurun_out/r02_stage_fp16.txt 2>&1; cat gpurun_out/r02_stage_fp16.txt | tail -16 )
( CALD_LIB=libcald_b200_bf16.so timeout 300 python tools/stage_error.py > gpurun_out/r02_stage_bf16.txt 2>&1; cat gpurun_out/r02_stage_bf16.txt | tail -16 )
( timeout 600 python tools/pool_parity.py frcnn gpurun_out/pool_engine_frcnn_fp16.npz > gpurun_out/r02_pool_frcnn_fp16.txt 2>&1; cat gpurun_out/r02_pool_frcnn_fp16.txt | tail -12 )
( CALD_LIB=libcald_b200_bf16.so timeout 600 python tools/pool_parity.py frcnn gpurun_out/pool_engine_frcnn_bf16.npz > gpurun_out/r02_pool_frcnn_bf16.txt 2>&1; cat gpurun_out/r02_pool_frcnn_bf16.txt | tail -12 )
echo "=== gpu suite (fp16 planes)"
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_pool.py 2>&1 | tail -30 > gpurun_out/r02_pytest1.txt; tail -30 gpurun_out/r02_pytest1.txt
for cfg in "fp16:" "bf16:CALD_LIB=libcald_b200_bf16.so" "resmma2:CALD_RESMMA_MAX_KB=2" "fuserpn:CALD_FUSE_RPN=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "=== bench $tag ($envs)"
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/r02_layers_$tag.tsv > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err
  tail -1 gpurun_out/r02_bench_$tag.json | cut -c1-260
  tail -3 gpurun_out/r02_bench_$tag.err
done
