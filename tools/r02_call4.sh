#!/bin/bash
# Round 2, GPU call 4: chunk = 64 images (all passes full size) A/B, ncu full-set captures of the dominant kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gpu suite (subset that exercises chunking)"
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_pool.py tests/test_gpu_api.py tests/test_gpu_fullsize.py tests/test_gpu_jpeg.py tests/test_gpu_retina.py -q 2>&1 | tail -40 > gpurun_out/r02d_pytest.txt; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r02d_pytest.txt | head -20
bench() { tag=$1; shift; echo "=== bench $tag: $*"; timeout 600 python bench.py "$@" > gpurun_out/r02d_bench_$tag.json 2> gpurun_out/r02d_bench_$tag.err; tail -1 gpurun_out/r02d_bench_$tag.json | cut -c1-300; tail -2 gpurun_out/r02d_bench_$tag.err; }
bench b64 --layers gpurun_out/r02d_layers_b64.tsv
bench b16 --batch 16 --views-per-pass 16 --steps 30 --warmup 5 --no-cpu-baseline
bench b128 --batch 128 --views-per-pass 128 --steps 6 --warmup 2 --no-cpu-baseline
echo "=== ncu full-set captures"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 40 -c 4 -o gpurun_out/r02d_pair \
    python bench.py --steps 1 --warmup 1 --only-value > gpurun_out/r02d_ncu_pair.log 2>&1; tail -2 gpurun_out/r02d_ncu_pair.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:igemm_tc_kernel -s 20 -c 6 -o gpurun_out/r02d_single \
    python bench.py --steps 1 --warmup 1 --only-value > gpurun_out/r02d_ncu_single.log 2>&1; tail -2 gpurun_out/r02d_ncu_single.log
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:roialign|view_stem_input|maxpool|pil_resample|consistency|nms_groups|topk_select' -c 12 -o gpurun_out/r02d_misc \
    python bench.py --steps 1 --warmup 0 --only-value > gpurun_out/r02d_ncu_misc.log 2>&1; tail -2 gpurun_out/r02d_ncu_misc.log
for f in pair single misc; do ncu -i gpurun_out/r02d_$f.ncu-rep --page raw --csv > gpurun_out/r02d_${f}_raw.csv 2>/dev/null; python tools/ncu_condense.py gpurun_out/r02d_${f}_raw.csv > gpurun_out/r02d_${f}_condensed.csv; done
ls -la gpurun_out/r02d_*
