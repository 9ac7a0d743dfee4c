#!/bin/bash
# shaft kernels compiled for 2 / 3 / 4 blocks per SM
cd "$(dirname "$0")/.."
for v in s2 s3 s4; do
  cp tools/variants/lib_$v.so fast_ray_tracer_b200/libfrt_b200.so
  echo "variant $v"
  python tools/ncu_frame.py 5 2>&1 | tail -2 | python -c "
import sys,re
for l in sys.stdin:
    m=re.search(r'frame \d+: ([0-9.]+) ms.*shadow_shaft.: ([0-9.]+)', l); print(m.group(1), 'ms frame, shaft', m.group(2)) if m else None"
done
cp tools/variants/lib_s3.so fast_ray_tracer_b200/libfrt_b200.so
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
