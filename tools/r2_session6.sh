#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/knn_debug.py > gpurun_out/s6_knn_debug.txt 2>&1
FRT_DEBUG_TIMING=1 python tools/e2e_probe2.py gen 2>&1 | tail -12 > gpurun_out/s6_e2e_probe.txt
cat gpurun_out/s6_knn_debug.txt | cut -c1-300; cat gpurun_out/s6_e2e_probe.txt
