#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest_last.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_last.txt
tail -15 gpurun_out/pytest_last.txt | cut -c1-250
