mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_b8.tsv > gpurun_out/bench_b8.log 2>&1; tail -1 gpurun_out/bench_b8.log | cut -c1-180
python bench.py --steps 6 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench_b16.log 2>&1; tail -1 gpurun_out/bench_b16.log | cut -c1-180
ncu --set full --clock-control none --import-source on -k regex:"roialign|topk_select|nms_groups" -s 8 -c 8 -f -o /tmp/prof_misc python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --workspace-gb 24 > gpurun_out/prof_misc.log 2>&1
ncu -i /tmp/prof_misc.ncu-rep --page raw --csv > gpurun_out/prof_misc_raw.csv 2>/dev/null
ncu -i /tmp/prof_misc.ncu-rep --page source --csv > gpurun_out/prof_misc_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
