#!/usr/bin/env python3
"""Render golden canvases with the UNMODIFIED reference (oracle/_ref/<scene>_ref) and store them as small
fixtures under tests/golden/ together with the flattened scene blob the CUDA path renders.

Run here (the reference tree must be present); the fixtures travel to the GPU box through git.
    python oracle/make_golden.py            # all entries of GOLDEN
    python oracle/make_golden.py cornell_exact_200
Each fixture <name>.npz holds: rgb (float32 linear canvas), srgb8 (uint8), meta (json: scene, size, spp, reference
seconds / threads / ray count from the wrapped intersect_world counter).
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import build_ref  # noqa: E402
from compare import read_canvas_dump, to_srgb8  # noqa: E402

REPO = Path(__file__).resolve().parent.parent
GOLD = REPO / "tests" / "golden"

# fixture name -> (scene, hsize, vsize, usteps, vsteps)   (0 = the scene's own value)
GOLDEN = {
    "reflect_refract": ("reflect_refract", 0, 0, 0, 0),
    "cornell_exact_200": ("cornell_exact", 200, 200, 4, 4),
    "cornell_exact_96_1spp": ("cornell_exact", 96, 96, 1, 1),
    "cornell_exact_800": ("cornell_exact", 800, 800, 4, 4),  # BASELINE.json configs[1] at full size (~7 min on 8 cores)
    "group_test": ("group_test", 0, 0, 0, 0),
    "csg_test": ("csg_test", 200, 200, 0, 0),
    "reflect_refract_test": ("reflect_refract_test", 0, 0, 0, 0),
    "checkered_torus": ("checkered_torus", 200, 200, 0, 0),
    "checkered_sphere": ("checkered_sphere", 200, 200, 0, 0),
    "checkered_cube": ("checkered_cube", 400, 200, 0, 0),
    "checkered_cylinder": ("checkered_cylinder", 200, 200, 0, 0),
    "align_check_plane": ("align_check_plane", 200, 200, 0, 0),
    "lens_test": ("lens_test", 300, 150, 0, 0),
    "shadow_glamour_shot": ("shadow_glamour_shot", 300, 120, 0, 0),
    "teapot": ("teapot", 200, 200, 0, 0),
    "bump_map_test": ("bump_map_test", 200, 200, 0, 0),
    "patterns_circle_light": ("patterns_circle_light", 0, 0, 0, 0),
    "dof_blur_240": ("dof_blur", 240, 120, 0, 0),
    "bounding_boxes_600": ("bounding_boxes", 600, 240, 0, 0),  # C3b: 6 x dragon.obj (141 K triangles, ~31 K groups)
    "texture_map_test": ("texture_map_test", 200, 200, 0, 0),
    # C4 stand-in: 86 K textured / bump-mapped OBJ triangles (smooth columns, glass windows), 10x10 area light, 2x2 CMJ
    "sibenik_surrogate_160": ("sibenik_surrogate", 160, 200, 2, 2),
    # photon-mapped: stochastic in the reference too, so two reference renders (seeds 1 and 2) are stored; their RMSE is
    # the noise floor the CUDA render is held against
    "cornell_gi_64": ("cornell_gi", 64, 64, 2, 2),
    "cornell_gi_caustics_48": ("cornell_gi_caustics", 48, 48, 1, 1),
    # C2 as benched, at a fixture-sized cache: every hit picks one of 64 cached sample sets with rand() twice
    # (light.c:194-198, renderer.c:915) -- stochastic in the reference itself, so two seeded renders are stored
    "cornell_cache64_200": ("cornell_cache64", 200, 200, 4, 4),
}
STOCHASTIC = {"cornell_gi_64", "cornell_gi_caustics_48", "dof_blur_240", "cornell_cache64_200"}
# fixtures whose scene blob is too large to commit (6 dragons = 52 MB): only the reference canvas is stored; the GPU test
# renders oracle/_ref/blobs/<scene>.frt, which travels to the GPU box with the snapshot
BLOB_STAYS_IN_REF = {"bounding_boxes_600", "sibenik_surrogate_160"}


def make(name: str):
    scene, hs, vs, us, vsteps = GOLDEN[name]
    env = {"FRT_COUNT_RAYS": "1"}
    if hs:
        env.update(FRT_REF_HSIZE=str(hs), FRT_REF_VSIZE=str(vs))
    if us:
        env.update(FRT_REF_USTEPS=str(us), FRT_REF_VSTEPS=str(vsteps))
    GOLD.mkdir(parents=True, exist_ok=True)
    rgb_b = None
    with tempfile.TemporaryDirectory() as td:
        dump = Path(td) / "canvas.bin"
        if name in STOCHASTIC:
            env["FRT_REF_SEED"] = "1"
        info = build_ref.run_reference(scene, dump, env)
        rgb = read_canvas_dump(dump)
        if name in STOCHASTIC:
            env2 = dict(env, FRT_REF_SEED="2")
            build_ref.run_reference(scene, dump, env2)
            rgb_b = read_canvas_dump(dump)
    if name not in BLOB_STAYS_IN_REF:
        blob = build_ref.dump_blob(scene, env, suffix=f"__{name}")
        shutil.move(str(blob), GOLD / f"{name}.frt")
    meta = {"scene": scene, "hsize": rgb.shape[1], "vsize": rgb.shape[0],
            "reference_seconds": float(info.get("FRT_RENDER_SECONDS", "nan")),
            "reference_threads": int(info.get("FRT_THREADS", "0")),
            "reference_rays": int(info.get("FRT_RAYS", "0")), "size_line": info.get("FRT_SIZE", "")}
    if "FRT_PHOTON_SECONDS" in info:
        meta["reference_photon_seconds"] = float(info["FRT_PHOTON_SECONDS"])
    extra = {"rgb_b": rgb_b.astype(np.float32)} if rgb_b is not None else {}
    np.savez_compressed(GOLD / f"{name}.npz", rgb=rgb.astype(np.float32), srgb8=to_srgb8(rgb).astype(np.uint8),
                        meta=json.dumps(meta), **extra)
    print(name, meta)


# PPM encoder fixtures: a float64 canvas the reference rendered and the bytes its own
# write_ppm_file (canvas.c:150-328) produced from that very canvas in the same run.
PPM_GOLDEN = {
    "ppm_cornell_exact_64": ("cornell_exact", 64, 64, 1, 1),           # dark frame: every channel maximum below 1
    "ppm_reflect_refract_100": ("reflect_refract", 100, 50, 0, 0),     # bright frame: the sum > sqrt(3) rescale is taken
    "ppm_checkered_sphere_64": ("checkered_sphere", 64, 64, 0, 0),
}


def make_ppm(name: str):
    scene, hs, vs, us, vsteps = PPM_GOLDEN[name]
    env = {"FRT_REF_HSIZE": str(hs), "FRT_REF_VSIZE": str(vs), "FRT_SKIP_PPM": "0"}
    if us:
        env.update(FRT_REF_USTEPS=str(us), FRT_REF_VSTEPS=str(vsteps))
    # output.file of the scene (the reference appends .ppm); yaml_parser/config.py:66-67 defaults it to /tmp/ray_tracer_out
    outs = [Path("/tmp/out_file.ppm"), Path("/tmp/ray_tracer_out.ppm")]
    for o in outs:
        if o.exists():
            o.unlink()
    with tempfile.TemporaryDirectory() as td:
        dump = Path(td) / "canvas.bin"
        build_ref.run_reference(scene, dump, env)
        raw = np.fromfile(dump, dtype=np.float64, offset=16)
        w, h = (int(x) for x in np.fromfile(dump, dtype=np.int64, count=2))
    rgb = raw.reshape(h, w, 3)
    out = next(o for o in outs if o.exists())
    ppm = np.frombuffer(out.read_bytes(), dtype=np.uint8)
    np.savez_compressed(GOLD / f"{name}.npz", rgb64=rgb, ppm=ppm, meta=json.dumps({"scene": scene, "hsize": w, "vsize": h, "use_scaling": True}))
    print(name, rgb.shape, len(ppm), "bytes, max rgb", rgb.reshape(-1, 3).max(axis=0))


if __name__ == "__main__":
    args = sys.argv[1:] or list(GOLDEN) + list(PPM_GOLDEN)
    for n in args:
        make_ppm(n) if n in PPM_GOLDEN else make(n)
