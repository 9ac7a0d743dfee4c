#!/usr/bin/env python3
"""Measure every BASELINE.json config on cuda:0 and, where its reference binary can run on this box (no OBJ assets),
the unmodified reference on the host cores.  Prints one JSON object per config (development aid; the numbers are copied
into BASELINE.md / profiles/)."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt  # noqa: E402
from fast_ray_tracer_b200.api import FRT_FLAG_COUNT_RAYS, FRT_FLAG_NO_PRUNE  # noqa: E402

G = REPO / "tests" / "golden"
REF = REPO / "oracle" / "_ref"
threads = len(os.sched_getaffinity(0))


def ref_run(binary, env):
    b = REF / binary
    if not b.exists():
        return None
    e = dict(os.environ, FRT_SKIP_PPM="1", FRT_REF_THREADS=str(threads), FRT_COUNT_RAYS="0")
    e.update(env)
    r = subprocess.run([str(b)], env=e, cwd=str(REF), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    out = {}
    for line in r.stdout.splitlines():
        if line.startswith("FRT_"):
            k, _, v = line.partition(" ")
            out[k] = v
    return out


def gpu(desc, reps=3, photons=None):
    with frt.Scene(desc) as sc:
        pm = None
        if photons:
            t0 = time.time()
            st = sc.trace_photons(3, bool(desc.config.gi_include_caustics), True, seed=1)
            pm = {"photon_pass_ms": 1e3 * (time.time() - t0), **st.extra}
        _, ref = sc.render(flags=FRT_FLAG_NO_PRUNE | FRT_FLAG_COUNT_RAYS, download=False) if not photons else (None, None)
        _, cnt = sc.render(flags=FRT_FLAG_COUNT_RAYS, download=False)
        ms = []
        for _ in range(reps):
            _, st = sc.render(download=False)
            ms.append(st.frame_ms)
    out = {"frame_ms": min(ms), "rays_traced": cnt.rays_total, "rays_shadow": cnt.rays_shadow, "rays_gather": cnt.rays_gather,
           "hits_shaded": cnt.hits_shaded, "shadow_nodes_per_ray": cnt.shadow_nodes / max(cnt.rays_shadow, 1),
           "algorithmic_flop_per_shadow_ray": cnt.light_flops / max(cnt.rays_shadow - cnt.shadow_deferred, 1),
           "deferred_fraction": cnt.shadow_deferred / max(cnt.rays_shadow, 1), "launches": cnt.kernel_launches,
           "shadow_rays_decided_per_hit_or_quadrant": cnt.extra["shadow_reasons"][0] / max(cnt.rays_shadow, 1)}
    if ref is not None:
        out["rays_reference_counted"] = ref.rays_total
    if pm:
        out.update(pm)
    return out


rows = {}
d = frt.SceneDesc.load(G / "reflect_refract.frt")
rows["C1 reflect_refract 400x200 1spp"] = gpu(d)
r = ref_run("reflect_refract_ref", {})
if r:
    rows["C1 reflect_refract 400x200 1spp"]["reference_frame_s"] = float(r["FRT_RENDER_SECONDS"])

d = frt.SceneDesc.load(G / "cornell_exact_200.frt")
d.set_resolution(800, 800)
rows["C2-exact cornell 800x800 4x4"] = gpu(d)
r = ref_run("cornell_exact_ref", {"FRT_REF_HSIZE": "200", "FRT_REF_VSIZE": "200"})
if r:
    rows["C2-exact cornell 800x800 4x4"]["reference_frame_s_scaled_from_200x200"] = 16 * float(r["FRT_RENDER_SECONDS"])

d = frt.SceneDesc.load(G / "teapot.frt")
d.set_resolution(400, 400)
rows["C3a teapot_low 400x400 1spp"] = gpu(d)

blob = REF / "blobs" / "bounding_boxes.frt"
if blob.exists():
    d = frt.SceneDesc.load(blob)
    rows["C3b bounding_boxes 1200x480 1spp (6 dragons)"] = gpu(d)

blob = REF / "blobs" / "sibenik_surrogate.frt"
if blob.exists():
    d = frt.SceneDesc.load(blob)  # 800 x 1000, 4 x 4 CMJ, 10 x 10 area light (one cached set), 85 K textured triangles
    key = "C4 stand-in (sibenik.obj is not in the reference tree) 800x1000 4x4, 10x10 area light"
    rows[key] = gpu(d, reps=2)
    r = ref_run("sibenik_surrogate_ref", {"FRT_REF_HSIZE": "80", "FRT_REF_VSIZE": "100", "FRT_REF_USTEPS": "2", "FRT_REF_VSTEPS": "2"})
    if r:
        rows[key]["reference_frame_s_scaled_from_80x100_2x2"] = 400 * float(r["FRT_RENDER_SECONDS"])

d = frt.SceneDesc.load(G / "cornell_gi_64.frt")
d.set_resolution(800, 800)
d.set_samples(4, 4)
d.config.gi_photon_count = 1000000
rows["C5 cornell GI 800x800 4x4, 1M photons, 8x8 gather"] = gpu(d, reps=1, photons=True)
r = ref_run("cornell_gi_ref", {"FRT_REF_HSIZE": "64", "FRT_REF_VSIZE": "64", "FRT_REF_USTEPS": "2", "FRT_REF_VSTEPS": "2"})
if r:
    rows["C5 cornell GI 800x800 4x4, 1M photons, 8x8 gather"]["reference_64x64_2x2_100Kphotons"] = {
        "frame_s": float(r["FRT_RENDER_SECONDS"]), "photon_pass_s": float(r.get("FRT_PHOTON_SECONDS", "nan"))}
d = frt.SceneDesc.load(G / "cornell_gi_64.frt")
rows["C5-small cornell GI 64x64 2x2, 100K photons"] = gpu(d, reps=2, photons=True)

print(json.dumps({"host_threads": threads, "configs": rows}, indent=1))
