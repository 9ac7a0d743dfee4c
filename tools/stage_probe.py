#!/usr/bin/env python3
"""Per-stage device times of a few frames of a scene blob (development aid): stage_probe.py blob [w h spp]."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt
d = frt.SceneDesc.load(sys.argv[1])
if len(sys.argv) > 3:
    d.set_resolution(int(sys.argv[2]), int(sys.argv[3]))
if len(sys.argv) > 4:
    d.set_samples(int(sys.argv[4]), int(sys.argv[4]))
with frt.Scene(d) as sc:
    for k in range(3):
        _, st = sc.render(download=False, seed=k, flags=256)
        print(f"frame {k}: {st.frame_ms:.2f} ms launches {st.kernel_launches} rays p/s/sh {st.rays_primary}/{st.rays_secondary}/{st.rays_shadow}",
              {a: round(b, 3) for a, b in st.extra["stage_ms"].items() if b > 0})
