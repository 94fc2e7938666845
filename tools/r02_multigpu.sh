#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): the product's sharding path (cfg5 cycle, strong scaling) and the weak-scaling
# configurations, one rank per GPU over NCCL.   usage: bash tools/r02_multigpu.sh N [pool]
N=${1:-2}; POOL=${2:-2048}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ "$N" -le 2 ]; then
  echo "=== RoIAlign / detection / pool parity tests (after the half2 lo-plane interpolation)"
  timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_golden.py tests/test_gpu_pool.py tests/test_gpu_select.py -q 2>&1 | tail -15 > gpurun_out/r02m_pytest.txt; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r02m_pytest.txt | head
  timeout 300 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --layers gpurun_out/r02m_layers_1gpu.tsv > gpurun_out/r02m_cfg2_1gpu.json 2> gpurun_out/r02m_cfg2_1gpu.err; tail -1 gpurun_out/r02m_cfg2_1gpu.json | cut -c1-200
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:roialign -c 2 python bench.py --steps 1 --warmup 0 --only-value --batch 16 2>&1 | grep -E "roialign|gpu__time" | head -6
fi
run() { tag=$1; shift; echo "=== $tag (N=$N): $*"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r02m_${tag}_${N}gpu.json 2> gpurun_out/r02m_${tag}_${N}gpu.err; tail -1 gpurun_out/r02m_${tag}_${N}gpu.json | cut -c1-260; grep -iE "error|Traceback" gpurun_out/r02m_${tag}_${N}gpu.err | head -3; }
run cfg5 --config cfg5 --pool $POOL --warmup 2
run cfg3 --config cfg3 --steps 6 --warmup 2 --quick
run cfg2 --steps 6 --warmup 2 --quick
[ "$N" -le 2 ] && run cfg4 --config cfg4 --steps 6 --warmup 2 --quick
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader | head -8
