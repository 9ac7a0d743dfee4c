#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/rep
timeout 600 python -m pytest tests/test_gpu_knn.py tests/test_gpu_gi.py -q -x 2>&1 | tail -4
FRT_KNN_DEBUG=1 timeout 300 python tools/gi_perf.py 400 4 2>&1 | tail -8 | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:^k_knn_cell -s 2 -c 1 -o /tmp/rep/gi_k_knn_cell python tools/gi_stage_probe.py 400 > gpurun_out/s24_ncu.log 2>&1
python tools/ncu_summary.py /tmp/rep/gi_k_knn_cell.ncu-rep 0 > gpurun_out/s24_k_knn_cell.txt 2>/dev/null
echo >> gpurun_out/s24_k_knn_cell.txt
python tools/ncu_lines.py /tmp/rep/gi_k_knn_cell.ncu-rep 24 >> gpurun_out/s24_k_knn_cell.txt 2>/dev/null
