#!/bin/bash
# Round 2, closing 8-GPU run on the final tree (gpurun --gpus 8): driver-style weak scaling, cfg-2, one rank per GPU over NCCL
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 6 --warmup 3 --quick > gpurun_out/r02y_mgpu_cfg2_8gpu.json 2> gpurun_out/r02y_mgpu_cfg2_8gpu.err
tail -1 gpurun_out/r02y_mgpu_cfg2_8gpu.json | cut -c1-300; grep -iE "error|Traceback" gpurun_out/r02y_mgpu_cfg2_8gpu.err | head -3
