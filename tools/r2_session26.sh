#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_knn.py tests/test_gpu_gi.py -q -x 2>&1 | tail -3
FRT_KNN_DEBUG=1 timeout 300 python tools/gi_perf.py 400 4 2>&1 | tail -4 | cut -c1-200
