#!/usr/bin/env python3
"""The bench workload (Cornell 800x800, 4x4, 65 535-set light cache rebuilt on the device) as a few plain production
frames -- the command the ncu captures under profiles/ are taken on (development aid; no timing claims are made here)."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt  # noqa: E402
from fast_ray_tracer_b200.lightcache import generate_area_light_caches  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
size = int(sys.argv[2]) if len(sys.argv) > 2 else 800
sets = int(sys.argv[3]) if len(sys.argv) > 3 else 65535
world = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # > 1: rank 0's share of the rows only (what one of N GPUs renders)
desc = frt.SceneDesc.load(REPO / "tests" / "golden" / "cornell_exact_200.frt")
desc.set_resolution(size, size)
desc.set_samples(4, 4)
if sets > 1:
    generate_area_light_caches(desc, sets, verify_sets=(0, sets - 1))
with frt.Scene(desc) as sc:
    for k in range(frames):
        _, st = sc.render(rank=0, world=world, download=False, seed=1000 + k, flags=256)
        print(f"frame {k}: {st.frame_ms:.3f} ms, k_shadow_f32 {st.light_ms:.3f} ms in {st.extra['shadow_ray_launches']} launches over "
              f"{st.extra['shadow_rays_traced']} rays; stages {st.extra['stage_ms']}")
