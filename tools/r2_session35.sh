#!/bin/bash
# k_extend compiled for 2 / 3 / 4 blocks per SM (shaft kernels at 4): Cornell frame, six dragons, C4 stand-in
cd "$(dirname "$0")/.."
for v in e2 e3 e4; do
  cp tools/variants/lib_$v.so fast_ray_tracer_b200/libfrt_b200.so
  echo "variant $v"
  python tools/ncu_frame.py 5 2>&1 | tail -1 | python -c "
import sys,re
for l in sys.stdin:
    m=re.search(r'frame \d+: ([0-9.]+) ms.*.extend.: ([0-9.]+).*shadow_shaft.: ([0-9.]+)', l); print(m.group(1), 'ms frame, extend', m.group(2), 'shaft', m.group(3)) if m else print(l[:200])"
  python tools/stage_probe.py oracle/_ref/blobs/bounding_boxes.frt 2>&1 | tail -1 | cut -c1-200
  python tools/stage_probe.py oracle/_ref/blobs/sibenik_surrogate.frt 400 500 4 2>&1 | tail -1 | cut -c1-200
done
cp tools/variants/lib_e2.so fast_ray_tracer_b200/libfrt_b200.so
