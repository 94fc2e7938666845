#!/bin/bash
# Round 2: compute-sanitizer memcheck over the kernels added this round (JPEG ingest, windowed RetinaNet NMS, on-device
# selection, batched Pillow kernels, upload pipeline).  Run under gpurun (1 GPU).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S="compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20"
run() { tag=$1; shift; echo "=== memcheck $tag"; timeout 700 $S python -m pytest "$@" -q -x 2>&1 | tail -25 > gpurun_out/r02s_$tag.txt; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r02s_$tag.txt | head -8; }
run jpeg tests/test_gpu_jpeg.py
run select "tests/test_gpu_select.py::test_reference_selection_fixture" "tests/test_gpu_select.py::test_pool_fixture_selection"
run retina_overflow "tests/test_gpu_api.py::test_retinanet_class_with_more_candidates_than_the_fast_list"
run score "tests/test_gpu_score.py::test_scores_match_oracle" "tests/test_gpu_score.py::test_batched_equals_single"
run augs "tests/test_gpu_augs_all.py" -k "ga or color"
