#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n1.err
tail -4 gpurun_out/r2_pytest_gpu.txt; tail -2 gpurun_out/r2_smoke.txt | cut -c1-300; tail -2 gpurun_out/r2_bench_n1.err
