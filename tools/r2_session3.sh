#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FRT_DEBUG_NODES=1 python tools/gpu_perf.py 800 4 1 cornell_exact_200 2 > gpurun_out/s3_entry_classes.txt 2>&1
python tools/ncu_frame.py 3 > gpurun_out/s3_frame_plain.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_shadow_entry -s 2 -c 1 -o gpurun_out/s3_k_shadow_entry python tools/ncu_frame.py 3 > gpurun_out/s3_ncu_full.log 2>&1
cat gpurun_out/s3_entry_classes.txt | cut -c1-200
