#!/bin/bash
# Round 2, GPU call 5: ncu full-set captures of the dominant kernels (reports stay on the box, CSV exports come back).
mkdir -p gpurun_out /tmp/ncu
export PYTHONUNBUFFERED=1
cap() { tag=$1; shift; timeout 500 ncu --set full --clock-control none --import-source on "$@" -o /tmp/ncu/$tag \
    python bench.py --steps 1 --warmup 1 --only-value --batch 16 > gpurun_out/r02e_ncu_$tag.log 2>&1; tail -1 gpurun_out/r02e_ncu_$tag.log
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > gpurun_out/r02e_${tag}_raw.csv 2>/dev/null
  python tools/ncu_condense.py gpurun_out/r02e_${tag}_raw.csv > gpurun_out/r02e_${tag}_condensed.csv; wc -l gpurun_out/r02e_${tag}_condensed.csv; }
# pair kernel: launches 40.. of the 64-view augmented pass (FPN / RPN 3x3 at P2..P5, box head)
cap pair -k regex:igemm_tc2 -s 50 -c 6
# one-CTA kernel: the short-K / residual launches of layer1-3 in the augmented pass
cap single -k regex:igemm_tc_kernel -s 34 -c 8
cap misc -k 'regex:roialign|view_stem_input|maxpool|pil_resample|consistency|nms_groups|topk_select|det_class_nms|cutout' -c 16
ls -la gpurun_out/r02e_*
