#!/bin/bash
# Round 2, closing multi-GPU sanity on the final tree (gpurun --gpus 2): the driver-style weak-scaling run and a short
# strong-scaling cycle through shard.get_uncertainty_sharded, one rank per GPU over NCCL
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { tag=$1; shift; echo "=== $tag: $*"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 "$@" > gpurun_out/r02y_mgpu_${tag}.json 2> gpurun_out/r02y_mgpu_${tag}.err; tail -1 gpurun_out/r02y_mgpu_${tag}.json | cut -c1-260; grep -iE "error|Traceback" gpurun_out/r02y_mgpu_${tag}.err | head -3; }
run cfg2 --steps 6 --warmup 3 --quick
run cfg5 --config cfg5 --pool 1024 --warmup 2
run reference --impl reference --steps 2 --warmup 1
