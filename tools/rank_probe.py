#!/usr/bin/env python3
"""Per-rank frame time of the row partition on ONE GPU (development aid): renders rank r of `world` for every r."""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt
import bench
desc = bench.load_workload(frt, "shipped", 800, 4)
with frt.Scene(desc) as sc:
    sc.render(download=False)
    for world in (1, 2, 4, 8):
        ms, wall = [], []
        for r in range(world):
            sc.render(rank=r, world=world, download=False)
            t0 = time.perf_counter()
            _, st = sc.render(rank=r, world=world, download=False, seed=3)
            wall.append(1e3 * (time.perf_counter() - t0))
            ms.append(st.frame_ms)
        print(f"world={world} frame_ms per rank: max {max(ms):.2f} min {min(ms):.2f} mean {sum(ms)/len(ms):.2f}  wall max {max(wall):.2f}  ideal {ms and (sum(ms)/1):.1f}")
