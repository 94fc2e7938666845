#!/bin/bash
# Round 2, GPU call 19: single-thread roles chosen with elect.sync (ptxas issues tcgen05.mma / TMA under it without the
# per-lane ELECT / BRA.U.ANY loop it wraps around `if (lane == 0)`): parity, then an A/B of the two builds on one box
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_golden.py -m gpu -q 2>&1 | tail -2 | tee gpurun_out/r02s_pytest.txt
for v in elect base elect base; do
  cp build_ab/libcald_b200_$v.so cald_b200/libcald_b200.so
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --layers gpurun_out/r02s_layers_$v.tsv > gpurun_out/r02s_bench_$v.json 2> gpurun_out/r02s_bench_$v.err
  echo "$v: $(python -c "import json;d=json.loads(open('gpurun_out/r02s_bench_$v.json').read().strip().splitlines()[-1]);print(round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_step'],2))")"
done
cp build_ab/libcald_b200_elect.so cald_b200/libcald_b200.so
python - <<'P'
import csv
a={r['layer']:r for r in csv.DictReader(open('gpurun_out/r02s_layers_elect.tsv'),delimiter='\t')}
b={r['layer']:r for r in csv.DictReader(open('gpurun_out/r02s_layers_base.tsv'),delimiter='\t')}
rows=sorted(a, key=lambda k:-float(b[k]['ms_total']) if k in b else 0)[:24]
for k in rows:
    if k in b: print("%-62s base %8.1f us  elect %8.1f us  %+5.1f %%" % (k[:62], float(b[k]['us_per_launch']), float(a[k]['us_per_launch']), 100*(float(a[k]['us_per_launch'])/float(b[k]['us_per_launch'])-1)))
P
