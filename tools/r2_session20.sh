#!/bin/bash
# leaf boxes + inserted groups over triangle runs: parity, then the two mesh scenes with and without the inserted groups
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s20_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s20_pytest.txt
tail -5 gpurun_out/s20_pytest.txt | cut -c1-200
for runs in 1 0; do
  echo "FRT_LEAF_RUNS=$runs"
  FRT_LEAF_RUNS=$runs timeout 300 python tools/dragons_perf.py 2>&1 | tail -3
  FRT_LEAF_RUNS=$runs timeout 300 python tools/sibenik_perf.py 400 500 4 2>&1 | tail -3
done 2>&1 | tee gpurun_out/s20_mesh.txt
