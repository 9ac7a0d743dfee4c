#!/usr/bin/env python3
"""Where an end-to-end step spends its time (development probe): scene creation with the light cache rebuilt on the
device, the frame, the canvas copy, destruction -- each fenced by a device synchronisation."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt  # noqa: E402
from fast_ray_tracer_b200.lightcache import expand_area_light_caches, generate_area_light_caches  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "gen"
desc = frt.SceneDesc.load(REPO / "tests" / "golden" / "cornell_exact_200.frt")
desc.set_resolution(800, 800)
desc.set_samples(4, 4)
if mode == "gen":
    generate_area_light_caches(desc, 65535, verify_sets=(0, 1, 4097, 32768, 65533, 65534))
elif mode == "gen0":
    generate_area_light_caches(desc, 65535)
else:
    expand_area_light_caches(desc, 65535)
    desc.pin()
out = torch.empty((800, 800, 4), dtype=torch.float64).pin_memory().numpy()


def now():
    torch.cuda.synchronize()
    return time.perf_counter()


for k in range(6):
    t0 = now()
    sc = frt.Scene(desc)
    t1 = now()
    _, st = sc.render(out=out, seed=k)
    t2 = now()
    sc.close()
    t3 = now()
    print(f"{mode} step {k}: create {1e3 * (t1 - t0):.2f} ms, render+download {1e3 * (t2 - t1):.2f} ms (device frame {st.frame_ms:.2f}, download {st.download_ms:.2f}), "
          f"destroy {1e3 * (t3 - t2):.2f} ms, total {1e3 * (t3 - t0):.2f} ms")
