#!/bin/bash
# k_light_final compiled for 4 / 5 / 6 blocks per SM
cd "$(dirname "$0")/.."
for v in f4 f5 f6; do
  cp tools/variants/lib_$v.so fast_ray_tracer_b200/libfrt_b200.so
  echo "variant $v"
  python tools/ncu_frame.py 5 2>&1 | tail -2 | python -c "
import sys,re
for l in sys.stdin:
    m=re.search(r'frame \d+: ([0-9.]+) ms.*light_final.: ([0-9.]+)', l); print(m.group(1), 'ms frame, light_final', m.group(2)) if m else print(l[:200])"
done
cp tools/variants/lib_f5.so fast_ray_tracer_b200/libfrt_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
