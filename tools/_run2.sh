mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2.log 2>&1; tail -1 gpurun_out/bench_r2.log | cut -c1-180
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --workspace-gb 24 > gpurun_out/ncu_r2.log 2>&1
