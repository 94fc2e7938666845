#!/bin/bash
# Round 2, GPU call 12: transposed-role kernel inside the engine -- stage / golden tests (stem + layer1 on it), then an
# A/B of the cfg-2 step on one box
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_detect.py tests/test_gpu_golden.py tests/test_gpu_models.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02l_pytest.txt
for t in 1 0 1 0; do
  CALD_TFORM=$t timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --layers gpurun_out/r02l_layers_tform$t.tsv > gpurun_out/r02l_bench_tform$t.json 2> gpurun_out/r02l_bench_tform$t.err
  echo "TFORM=$t: $(python -c "import json;d=json.loads(open('gpurun_out/r02l_bench_tform$t.json').read().strip().splitlines()[-1]);print(round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4))")"
done
grep -E "k7 cin3|k3 cin64" gpurun_out/r02l_layers_tform1.tsv gpurun_out/r02l_layers_tform0.tsv
