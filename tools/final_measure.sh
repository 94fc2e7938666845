#!/bin/bash
# Round-end measurement bundle (run under gpurun, 1 GPU): GPU tests, the two bench configurations, per-layer tables,
# the ncu launch list of one step, the DRAM traffic of every conv launch of that step and a full-set capture of the
# CTA-pair conv kernel on the FPN / RPN 3x3 layers of the augmented pass.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -x 2>&1 | tail -4 > gpurun_out/final_pytest_conv.txt; cat gpurun_out/final_pytest_conv.txt
grep -q "failed\|error" gpurun_out/final_pytest_conv.txt && { echo "conv tests failed - stopping"; exit 1; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
python bench.py --layers gpurun_out/final_layers_frcnn.tsv > gpurun_out/final_bench_frcnn.json 2> gpurun_out/final_bench_frcnn.err; tail -1 gpurun_out/final_bench_frcnn.json | cut -c1-160
python bench.py --model retinanet --no-cpu-baseline --layers gpurun_out/final_layers_retina.tsv > gpurun_out/final_bench_retina.json 2> gpurun_out/final_bench_retina.err; tail -1 gpurun_out/final_bench_retina.json | cut -c1-160
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workspace-gb 48"
ncu --metrics gpu__time_duration.sum --clock-control none -s 256 -c 256 --csv --log-file gpurun_out/final_launches_step.csv $CMD > gpurun_out/final_ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:igemm_tc -s 142 -c 142 --csv --log-file gpurun_out/final_igemm_dram_step.csv $CMD > gpurun_out/final_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 142 -c 14 -f -o /tmp/prof_pair $CMD > gpurun_out/final_ncu3.log 2>&1
ncu -i /tmp/prof_pair.ncu-rep --page raw --csv > /tmp/prof_pair_raw.csv 2>/dev/null
python tools/ncu_condense.py /tmp/prof_pair_raw.csv > gpurun_out/final_pair_full_capture.csv
ls -la gpurun_out | grep final
