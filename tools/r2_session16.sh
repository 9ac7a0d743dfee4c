#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/gpu_perf.py 800 4 2 cornell_exact_200 0,512,2,18 2>&1 | cut -c1-330 > gpurun_out/s16_perf.txt
python tools/ncu_frame.py 3 | tail -1 | cut -c1-420 >> gpurun_out/s16_perf.txt
timeout 1700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gi.py tests/test_gpu_dropin.py -q -x > gpurun_out/s16_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/s16_pytest.txt
tail -4 gpurun_out/s16_pytest.txt; cat gpurun_out/s16_perf.txt
