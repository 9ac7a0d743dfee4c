#!/bin/bash
cd "$(dirname "$0")/.."
FRT_KNN_SORT=0 python tools/gi_stage_probe.py 400 | tail -1 | cut -c1-400
FRT_KNN_SORT=1 python tools/gi_stage_probe.py 400 | tail -1 | cut -c1-400
python -m pytest tests/test_gpu_gi.py tests/test_gpu_knn.py -q 2>&1 | tail -3
