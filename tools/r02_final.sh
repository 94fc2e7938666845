#!/bin/bash
# Round 2, final GPU call: whole GPU suite, sanitizer over the new kernel, the BASELINE configurations through bench.py,
# the reference arm, the ncu launch list (time + DRAM bytes per launch) and a full-set capture of igemm_t_kernel.
mkdir -p gpurun_out /tmp/ncu
export PYTHONUNBUFFERED=1
echo "=== gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$|Warning|warnings.warn|cos = |sin = " > gpurun_out/r02z_pytest_full.txt; tail -3 gpurun_out/r02z_pytest_full.txt
grep -E "full-size pool|edge shapes|pool:|one-seed" gpurun_out/r02z_pytest_full.txt | head
echo "=== sanitizer (memcheck, racecheck) over the transposed-role kernel"
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 10 python -m pytest tests/test_gpu_conv.py -q -x -k "transposed and not case4" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" | head -4 | sed "s/^/[$tool] /" | tee -a gpurun_out/r02z_sanitizer.txt
done
bench() { tag=$1; shift; echo "=== bench $tag: $*"; ( time timeout 700 python bench.py "$@" > gpurun_out/r02z_bench_$tag.json 2> gpurun_out/r02z_bench_$tag.err ) 2>&1 | grep real; tail -1 gpurun_out/r02z_bench_$tag.json | cut -c1-330; tail -1 gpurun_out/r02z_bench_$tag.err | cut -c1-200; }
bench cfg2 --gpus 1 --steps 20 --warmup 5 --layers gpurun_out/r02z_layers_cfg2.tsv
bench reference --impl reference --gpus 1 --steps 12 --warmup 3
bench cfg3 --config cfg3 --steps 10 --warmup 3 --cpu-images 2 --layers gpurun_out/r02z_layers_cfg3.tsv
bench cfg4 --config cfg4 --steps 10 --warmup 3 --cpu-images 3 --layers gpurun_out/r02z_layers_cfg4.tsv
bench cfg5 --config cfg5 --pool 2048 --warmup 3
echo "=== ncu launch list (time + DRAM bytes per launch), one 16-image step"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 420 \
    --csv --log-file gpurun_out/r02z_launches_step.csv python bench.py --steps 1 --warmup 1 --only-value --batch 16 > gpurun_out/r02z_ncu.log 2>&1
tail -2 gpurun_out/r02z_ncu.log; wc -l gpurun_out/r02z_launches_step.csv
echo "=== ncu full set: igemm_t_kernel (stem, layer1 3x3)"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:igemm_t_kernel -s 4 -c 4 -o /tmp/ncu/tform \
    python bench.py --steps 1 --warmup 1 --only-value --batch 16 > gpurun_out/r02z_ncu_tform.log 2>&1; tail -1 gpurun_out/r02z_ncu_tform.log
ncu -i /tmp/ncu/tform.ncu-rep --page raw --csv > gpurun_out/r02z_tform_raw.csv 2>/dev/null
python tools/ncu_condense.py gpurun_out/r02z_tform_raw.csv > gpurun_out/r02z_tform_condensed.csv; wc -l gpurun_out/r02z_tform_condensed.csv
