#!/bin/bash
# Round 2, GPU call 11: the transposed-role kernel -- parity tests, then the layer micro-benchmark
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -s -k "transposed or pair" 2>&1 | grep -vE "^\s*$|Warning|warnings.warn" | tail -30 | tee gpurun_out/r02k_pytest_tform.txt
echo "=== micro"
timeout 200 python tools/tform_micro.py 16 2>&1 | tail -8 | tee gpurun_out/r02k_tform_micro.txt
