#!/bin/bash
# k_knn_cell: requests sorted by cell, a lane per request, candidates shared through shared memory
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn.py tests/test_gpu_gi.py -q -x 2>&1 | tail -15
for m in 1 0; do echo "FRT_KNN_MODE=$m"; FRT_KNN_MODE=$m timeout 300 python tools/gi_perf.py 400 4 2>&1 | tail -2; done | tee gpurun_out/s21_gi.txt
FRT_DEBUG_TIMING=1 timeout 300 python tools/gi_stage_probe.py 2>&1 | tail -12 | cut -c1-400
