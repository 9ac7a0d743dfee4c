#!/bin/bash
# k_shadow_mesh variants: HEAD, inline triangle + trimmed state at 6 / 5 / 4 blocks per SM
cd "$(dirname "$0")/.."
for v in head m6 m5 m4; do
  cp tools/variants/lib_$v.so fast_ray_tracer_b200/libfrt_b200.so
  echo "variant $v"
  timeout 300 python tools/dragons_perf.py 2>&1 | sed -n 2p | cut -c1-60
  timeout 300 python tools/sibenik_perf.py 400 500 4 2>&1 | sed -n 2p | cut -c1-60
done
cp tools/variants/lib_m6.so fast_ray_tracer_b200/libfrt_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "teapot or bounding or sibenik or group" 2>&1 | tail -2
