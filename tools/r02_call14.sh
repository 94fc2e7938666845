#!/bin/bash
# Round 2, GPU call 14: transposed-role kernel with the 16x256b epilogue -- parity, micro-benchmark, stem on / off in the step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -s -k "transposed" 2>&1 | grep -vE "^\s*$|Warning|warnings.warn" | tail -12 | tee gpurun_out/r02o_pytest_tform.txt
echo "=== micro"
timeout 200 python tools/tform_micro.py 16 2>&1 | tail -7 | tee gpurun_out/r02o_tform_micro.txt
echo "=== engine tests with the stem on the kernel too"
CALD_TFORM_STEM=1 timeout 300 python -m pytest tests/test_gpu_detect.py tests/test_gpu_golden.py -m gpu -q 2>&1 | tail -3
for t in 1 0 1 0; do
  CALD_TFORM_STEM=$t timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --layers gpurun_out/r02o_layers_stem$t.tsv > gpurun_out/r02o_bench_stem$t.json 2> gpurun_out/r02o_bench_stem$t.err
  echo "TFORM_STEM=$t: $(python -c "import json;d=json.loads(open('gpurun_out/r02o_bench_stem$t.json').read().strip().splitlines()[-1]);print(round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4))")"
done
grep -E "k7 cin3|k3 cin64" gpurun_out/r02o_layers_stem1.tsv gpurun_out/r02o_layers_stem0.tsv
