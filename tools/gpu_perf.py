#!/usr/bin/env python3
"""Frame-time probe: cornell_box exact at a given size (development aid)."""
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 800
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
blob = sys.argv[4] if len(sys.argv) > 4 else "cornell_exact_200"
flag_list = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0, 2, 2 | 16, 8]
desc = frt.SceneDesc.load(REPO / "tests" / "golden" / f"{blob}.frt")
desc.set_resolution(size, size)
desc.set_samples(spp, spp)
with frt.Scene(desc) as sc:
    for flags in flag_list:
        for r in range(reps):
            t0 = time.time()
            _, st = sc.render(flags=flags, download=False)
            wall = time.time() - t0
            tot = st.rays_total
            print(f"flags={flags} frame_ms={st.frame_ms:.2f} light_ms={st.light_ms:.2f} wall_ms={wall*1e3:.1f} launches={st.kernel_launches} "
                  f"rays p/s/sh={st.rays_primary}/{st.rays_secondary}/{st.rays_shadow} Mrays/s={tot/st.frame_ms/1e3:.1f} "
                  f"hits={st.hits_shaded} nodes/shadowray={st.shadow_nodes/max(st.rays_shadow,1):.2f} deferred={st.shadow_deferred} mismatch={st.shadow_mismatch} reasons={st.extra.get('shadow_reasons')}")
