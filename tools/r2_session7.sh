#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/s7_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/s7_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
echo "bench rc=$?" >> gpurun_out/s7_bench.err
tail -8 gpurun_out/s7_pytest.txt | cut -c1-250; tail -5 gpurun_out/s7_bench.err
