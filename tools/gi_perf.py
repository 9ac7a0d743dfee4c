#!/usr/bin/env python3
"""C5 probe: cornell_box with the shipped GI configuration (global map, 8x8 final gather, kNN 200, r 0.1) at a given
size / photon count (development aid)."""
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 200
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
photons = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
desc = frt.SceneDesc.load(REPO / "tests" / "golden" / "cornell_gi_64.frt")
desc.set_resolution(size, size)
desc.set_samples(spp, spp)
desc.config.gi_photon_count = photons
with frt.Scene(desc) as sc:
    for rep in range(2):
        t0 = time.time()
        st = sc.trace_photons(3, False, True, seed=rep)
        t1 = time.time()
        print(f"photon pass {1e3*(t1-t0):.1f} ms emitted={st.extra['rays_photon']} stored={st.extra['photons_stored']}")
        _, rs = sc.render(download=False, seed=rep)
        print(f"frame_ms={rs.frame_ms:.1f} wall={time.time()-t1:.2f}s gather_rays={rs.rays_gather} hits={rs.hits_shaded} launches={rs.kernel_launches}")
