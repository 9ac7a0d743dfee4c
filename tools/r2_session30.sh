#!/bin/bash
# one light stage for the hits of every level: parity, then the bench frame on one GPU and rank 0's share of eight
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s30_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s30_pytest.txt
tail -3 gpurun_out/s30_pytest.txt | cut -c1-200
for m in 1 0; do
  echo "FRT_MERGE_LEVELS=$m"
  FRT_MERGE_LEVELS=$m python tools/ncu_frame.py 4 800 65535 1 2>&1 | tail -2 | cut -c1-120
  FRT_MERGE_LEVELS=$m python tools/ncu_frame.py 4 800 65535 8 2>&1 | tail -2 | cut -c1-120
done
FRT_MERGE_LEVELS=1 python tools/dragons_perf.py 2>&1 | tail -2 | cut -c1-100
FRT_MERGE_LEVELS=0 python tools/dragons_perf.py 2>&1 | tail -2 | cut -c1-100
