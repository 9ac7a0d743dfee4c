#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in gen gen0 upload; do python tools/e2e_probe2.py $m; done > gpurun_out/s2_e2e_probe.txt 2>&1
FRT_ENTRY_KERNEL=0 python tools/ncu_frame.py 3 > gpurun_out/s2_frame_item.txt 2>&1
FRT_ENTRY_KERNEL=1 python tools/ncu_frame.py 3 > gpurun_out/s2_frame_entry.txt 2>&1
for b in 16 32 128; do echo "blocks x$b"; FRT_ENTRY_BLOCKS=$((148*b)) python tools/ncu_frame.py 3 | tail -1 | cut -c1-120; done > gpurun_out/s2_frame_entry_blocks.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/s2_pytest.txt
tail -5 gpurun_out/s2_pytest.txt; cat gpurun_out/s2_e2e_probe.txt; cut -c1-130 gpurun_out/s2_frame_item.txt gpurun_out/s2_frame_entry.txt; cat gpurun_out/s2_frame_entry_blocks.txt
