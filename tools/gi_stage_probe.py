#!/usr/bin/env python3
"""C5 (Cornell + 1 M-photon global map + 8x8 final gather) at a given size: per-stage device times of one frame."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt
size = int(sys.argv[1]) if len(sys.argv) > 1 else 400
photons = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
d = frt.SceneDesc.load(REPO / "tests/golden/cornell_gi_64.frt")
d.set_resolution(size, size)
d.set_samples(4, 4)
d.config.gi_photon_count = photons
with frt.Scene(d) as sc:
    sc.trace_photons(3, False, True, seed=7)
    for k in range(2):
        _, st = sc.render(download=False, seed=k, flags=256)
        print(f"frame {k}: {st.frame_ms:.1f} ms gather rays {st.rays_gather} stages", {a: round(b, 2) for a, b in st.extra["stage_ms"].items() if b > 0})
