#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FRT_DEBUG_TIMING=1 python tools/e2e_probe2.py gen > gpurun_out/s5_e2e_probe.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_knn.py tests/test_gpu_gi.py tests/test_gpu_lightgen.py tests/test_gpu_multi.py tests/test_gpu_dropin.py -q > gpurun_out/s5_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/s5_pytest.txt
tail -30 gpurun_out/s5_pytest.txt | cut -c1-200; tail -60 gpurun_out/s5_e2e_probe.txt
