#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/rep
python tools/gi_stage_probe.py 160 > gpurun_out/s23_gi_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_knn_cell -s 1 -c 1 -o /tmp/rep/gi_k_knn_cell python tools/gi_stage_probe.py 160 > gpurun_out/s23_ncu.log 2>&1
python tools/ncu_summary.py /tmp/rep/gi_k_knn_cell.ncu-rep 0 > gpurun_out/s23_k_knn_cell.txt 2>/dev/null
echo >> gpurun_out/s23_k_knn_cell.txt
python tools/ncu_lines.py /tmp/rep/gi_k_knn_cell.ncu-rep 40 >> gpurun_out/s23_k_knn_cell.txt 2>/dev/null
tail -3 gpurun_out/s23_gi_plain.txt | cut -c1-300
