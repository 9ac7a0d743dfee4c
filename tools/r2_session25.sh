#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s25_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s25_pytest.txt
tail -4 gpurun_out/s25_pytest.txt | cut -c1-200
FRT_KNN_DEBUG=1 timeout 300 python tools/gi_perf.py 400 4 2>&1 | tail -4 | cut -c1-200
timeout 300 python tools/gi_stage_probe.py 800 2>&1 | tail -1 | cut -c1-400
