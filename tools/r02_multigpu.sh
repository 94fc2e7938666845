#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): the product's sharding path (cfg5 cycle, strong scaling) and the weak-scaling
# configurations, one rank per GPU over NCCL.   usage: bash tools/r02_multigpu.sh N [pool]
N=${1:-2}; POOL=${2:-2048}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { tag=$1; shift; echo "=== $tag (N=$N): $*"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r02m_${tag}_${N}gpu.json 2> gpurun_out/r02m_${tag}_${N}gpu.err; tail -1 gpurun_out/r02m_${tag}_${N}gpu.json | cut -c1-260; grep -iE "error|Traceback" gpurun_out/r02m_${tag}_${N}gpu.err | head -3; }
run cfg5 --config cfg5 --pool $POOL --warmup 2
run cfg3 --config cfg3 --steps 6 --warmup 2 --quick
run cfg2 --steps 6 --warmup 2 --quick
[ "$N" -le 2 ] && run cfg4 --config cfg4 --steps 6 --warmup 2 --quick
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader | head -8
