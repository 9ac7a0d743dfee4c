#!/usr/bin/env python3
"""Render every golden fixture on cuda:0 through the C ABI and print the parity report (development aid)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))
import fast_ray_tracer_b200 as frt  # noqa: E402
from compare import parity_report  # noqa: E402

names = sys.argv[1:] or sorted(p.stem for p in (REPO / "tests" / "golden").glob("*.npz"))
for name in names:
    z = np.load(REPO / "tests" / "golden" / f"{name}.npz")
    desc = frt.SceneDesc.load(REPO / "tests" / "golden" / f"{name}.frt")
    t0 = time.time()
    try:
        canvas, st = frt.render_multi(desc, flags=2)
    except frt.FrtError as e:
        print(name, "ERROR", e)
        continue
    rep = parity_report(canvas[..., :3], z["rgb"].astype(np.float64))
    meta = json.loads(str(z["meta"]))
    print(name, json.dumps(rep), f"frame_ms={st.frame_ms:.2f} light_ms={st.light_ms:.2f} wall={time.time()-t0:.2f}s",
          f"rays p/s/sh={st.rays_primary}/{st.rays_secondary}/{st.rays_shadow} ref_rays={meta['reference_rays']} ref_s={meta['reference_seconds']:.2f}")
    np.save(REPO / "gpurun_out" / f"{name}_gpu.npy", canvas[..., :3].astype(np.float32))
