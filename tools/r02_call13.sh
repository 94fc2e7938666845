#!/bin/bash
# Round 2, GPU call 13: narrow identity MMAs for the residual k-blocks -- conv parity tests, then an A/B of the step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_golden.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02n_pytest.txt
for t in 1 0 1 0; do
  CALD_RES_NARROW=$t timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --layers gpurun_out/r02n_layers_narrow$t.tsv > gpurun_out/r02n_bench_narrow$t.json 2> gpurun_out/r02n_bench_narrow$t.err
  echo "RES_NARROW=$t: $(python -c "import json;d=json.loads(open('gpurun_out/r02n_bench_narrow$t.json').read().strip().splitlines()[-1]);print(round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4))")"
done
grep -E "resmma" gpurun_out/r02n_layers_narrow1.tsv gpurun_out/r02n_layers_narrow0.tsv
