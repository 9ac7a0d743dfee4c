#!/usr/bin/env python3
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "oracle"))
import fast_ray_tracer_b200 as frt
import pm_ref
desc = frt.SceneDesc.load(REPO / "tests/golden/cornell_gi_64.frt")
cfg = desc.config
rng = np.random.default_rng(11)
with frt.Scene(desc) as sc:
    sc.trace_photons(3, False, True, seed=5)
    rec = sc.photons_export(1)
    n = rec.shape[1]
    pos, power = rec[0, :, :3], rec[1, :, :3]
    bits = np.ascontiguousarray(rec[0, :, 3]).view(np.uint32)
    theta, phi = (bits & 255).astype(np.uint8), ((bits >> 8) & 255).astype(np.uint8)
    pick = rng.integers(0, n, 3000)
    q1 = pos[pick].astype(np.float64) + rng.normal(0, 0.01, (3000, 3))
    q2 = rng.uniform(-1.5, 1.5, (1000, 3))
    qpos = np.concatenate([q1, q2]).astype(np.float32).astype(np.float64)
    qn = rng.standard_normal(qpos.shape); qn /= np.linalg.norm(qn, axis=1, keepdims=True); qn = qn.astype(np.float32).astype(np.float64)
    irr, found = sc.photons_estimate(1, qpos, qn)
ref, rfound = pm_ref.estimate(pos, power, theta, phi, qpos, qn, cfg.gi_irradiance_estimate_radius, cfg.gi_irradiance_estimate_num, cfg.gi_irradiance_estimate_cone_filter_k)
R = cfg.gi_irradiance_estimate_radius
print("n photons", n, "radius", R, "num", cfg.gi_irradiance_estimate_num, "pos range", pos.min(0), pos.max(0))
bad = np.nonzero(found != rfound)[0]
for i in bad:
    d2 = ((pos.astype(np.float64) - qpos[i]) ** 2).sum(1)
    inside = d2 < R * R
    print(i, "device", found[i], "ref", rfound[i], "brute", inside.sum(), "q", qpos[i], "nearest outside/inside margin", np.sort(np.abs(np.sqrt(d2) - R))[:3])
    if inside.sum() != found[i]:
        miss = np.nonzero(inside)[0]
        print("   inside photons x range", pos[miss].min(0), pos[miss].max(0))
