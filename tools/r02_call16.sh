#!/bin/bash
# Round 2, GPU call 16: transposed-role kernel with a ring of four half stages -- conv parity, the pool / golden / detect
# tests (the cross-term partial sums of a k-block are now added in another order), micro-benchmark, step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -s -k "transposed" 2>&1 | grep -E "tform|passed|failed" | tee gpurun_out/r02p_pytest_tform.txt
timeout 200 python tools/tform_micro.py 16 2>&1 | tail -7 | tee gpurun_out/r02p_tform_micro.txt
timeout 900 python -m pytest tests/test_gpu_pool.py tests/test_gpu_golden.py tests/test_gpu_detect.py tests/test_gpu_fullsize.py tests/test_gpu_score.py tests/test_gpu_configs.py tests/test_gpu_retina.py -m gpu -q -s 2>&1 | grep -E "pool|one-seed|passed|failed|^FAILED|^E " | head -30 | tee gpurun_out/r02p_pytest_parity.txt
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --layers gpurun_out/r02p_layers.tsv > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
python -c "import json;d=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1]);print(round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4))"
grep -E "k7 cin3|k3 cin64" gpurun_out/r02p_layers.tsv
