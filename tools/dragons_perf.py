#!/usr/bin/env python3
"""bounding_boxes.yml (6 dragons) frame probe (development aid)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt
d = frt.SceneDesc.load(REPO / "oracle" / "_ref" / "blobs" / "bounding_boxes.frt")
with frt.Scene(d) as sc:
    for f in (0, 0, 2):
        _, st = sc.render(flags=f, download=False)
        print(f, f"frame_ms={st.frame_ms:.1f} light_ms={st.light_ms:.2f} p/s/sh={st.rays_primary}/{st.rays_secondary}/{st.rays_shadow} deferred={st.shadow_deferred} nodes/shadow={st.shadow_nodes/max(st.rays_shadow,1):.1f} launches={st.kernel_launches}")
