#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FRT_ENTRY_KERNEL=1 FRT_DEBUG_NODES=1 python tools/gpu_perf.py 800 4 1 cornell_exact_200 2,18 2>&1 | grep -v "undecided at" > gpurun_out/s9_entry_classes.txt
FRT_ENTRY_KERNEL=0 python tools/ncu_frame.py 3 > gpurun_out/s9_frame_item.txt 2>&1
FRT_ENTRY_KERNEL=1 python tools/ncu_frame.py 3 > gpurun_out/s9_frame_entry.txt 2>&1
timeout 1700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lightgen.py tests/test_gpu_knn.py -q -x > gpurun_out/s9_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/s9_pytest.txt
FRT_DEBUG_TIMING=1 python tools/e2e_probe2.py gen 2>&1 | tail -10 > gpurun_out/s9_e2e_probe.txt
tail -4 gpurun_out/s9_pytest.txt; cut -c1-200 gpurun_out/s9_entry_classes.txt; tail -1 gpurun_out/s9_frame_item.txt | cut -c1-420; tail -1 gpurun_out/s9_frame_entry.txt | cut -c1-420; cat gpurun_out/s9_e2e_probe.txt
