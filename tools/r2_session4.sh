#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FRT_DEBUG_NODES=1 python tools/gpu_perf.py 800 4 1 cornell_exact_200 2,18 > gpurun_out/s4_entry_classes.txt 2>&1
FRT_ENTRY_KERNEL=0 python tools/ncu_frame.py 3 > gpurun_out/s4_frame_item.txt 2>&1
FRT_ENTRY_KERNEL=1 python tools/ncu_frame.py 3 > gpurun_out/s4_frame_entry.txt 2>&1
timeout 1700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gi.py -q -x > gpurun_out/s4_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/s4_pytest.txt
tail -5 gpurun_out/s4_pytest.txt; cut -c1-220 gpurun_out/s4_entry_classes.txt; cut -c1-400 gpurun_out/s4_frame_item.txt gpurun_out/s4_frame_entry.txt
