#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/gi_stage_probe.py 160 > gpurun_out/s14_plain.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_knn -s 1 -c 1 -o gpurun_out/s14_k_knn python tools/gi_stage_probe.py 160 > gpurun_out/s14_ncu.log 2>&1
tail -2 gpurun_out/s14_ncu.log; cat gpurun_out/s14_plain.txt | cut -c1-300
