#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/ncu_frame.py 3 > gpurun_out/s10_frame_plain.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_shadow_f32 -s 2 -c 1 -o gpurun_out/s10_k_shadow_f32 python tools/ncu_frame.py 3 > gpurun_out/s10_ncu_full.log 2>&1
tail -2 gpurun_out/s10_ncu_full.log
