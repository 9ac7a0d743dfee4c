#!/usr/bin/env python3
"""C4 stand-in frame probe: sibenik surrogate at a given size (development aid)."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt  # noqa: E402

w = int(sys.argv[1]) if len(sys.argv) > 1 else 160
h = int(sys.argv[2]) if len(sys.argv) > 2 else 200
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = frt.SceneDesc.load(REPO / "oracle" / "_ref" / "blobs" / "sibenik_surrogate.frt")
d.set_resolution(w, h)
d.set_samples(spp, spp)
with frt.Scene(d) as sc:
    for f in (0, 0, 2):
        _, st = sc.render(flags=f, download=False)
        print(f, f"frame_ms={st.frame_ms:.1f} light_ms={st.light_ms:.2f} p/s/sh={st.rays_primary}/{st.rays_secondary}/{st.rays_shadow} "
                 f"deferred={st.shadow_deferred} nodes/shadow={st.shadow_nodes/max(st.rays_shadow,1):.1f} launches={st.kernel_launches}")
