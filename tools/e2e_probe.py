#!/usr/bin/env python3
"""End-to-end step probe: scene create / render / destroy wall times (development aid)."""
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt  # noqa: E402

from fast_ray_tracer_b200.lightcache import expand_area_light_caches  # noqa: E402

desc = frt.SceneDesc.load(REPO / "tests" / "golden" / "cornell_exact_200.frt")
desc.set_resolution(800, 800)
desc.set_samples(4, 4)
if len(sys.argv) > 1 and int(sys.argv[1]) > 1:
    expand_area_light_caches(desc, int(sys.argv[1]))
if len(sys.argv) > 4 and sys.argv[4] == "pin":
    desc.pin()
if len(sys.argv) > 2 and sys.argv[2] == "nvml":  # bench.py's clock sampler next to the loop
    sys.path.insert(0, str(REPO))
    from bench import ClockSampler

    cs = ClockSampler(0)
    cs.start()
out = np.empty((800, 800, 4), dtype=np.float64)
keep = frt.Scene(desc)
keep.render(download=False)
for k in range(int(sys.argv[3]) if len(sys.argv) > 3 else 4):
    t0 = time.perf_counter()
    sc = frt.Scene(desc)
    t1 = time.perf_counter()
    _, st = sc.render(out=out, seed=k)
    t2 = time.perf_counter()
    sc.close()
    t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms  render+download {1e3*(t2-t1):.1f} ms (frame {st.frame_ms:.1f}, download {st.download_ms:.1f})  destroy {1e3*(t3-t2):.1f} ms")
