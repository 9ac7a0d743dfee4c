#!/usr/bin/env python3
"""Build the reference-backed test artefacts under oracle/_ref/ (TEST INFRASTRUCTURE, not product).

The reference (gbordelon/fast_ray_tracer, C11) compiles from its own source files with plain gcc, so the
strongest oracle is the reference itself.  This script compiles the sources WHERE THEY LIE under
/root/reference (nothing is copied into the repository), writes only into oracle/_ref/ (git-ignored, but it
travels to the GPU box with the gpurun snapshot) and produces, per scene:

    oracle/_ref/<scene>_ref    the unmodified reference program (pthread CPU renderer), wrapped by
                               oracle/ref_hooks.c for timing / ray counting / raw canvas dumps
    oracle/_ref/<scene>_b200   the same generated main.c + the reference's host-side scene construction, with
                               renderer.c and photon_tracer.c replaced by fast_ray_tracer_b200/csrc/frt_shim.c
                               and linked against libfrt_b200.so  (the drop-in build of INTEGRATION.md)
    oracle/_ref/blobs/<scene>.frt   the flattened scene (frt_scene_save), made by running <scene>_b200 with
                               FRT_DUMP_ONLY=1 -- no GPU needed

Accommodations, all harness-side: -std=gnu11 (drand48 / M_PI), core_select.c excluded
(mach-only), a stub png.h (oracle/png_stub).  The reference's own build system is not run.
"""
from __future__ import annotations

import argparse
import copy
import os
import shutil
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("FRT_REFERENCE_ROOT", "/root/reference"))
OUT = REPO / "oracle" / "_ref"
OBJ = OUT / "obj"
GEN = OUT / "gen"
BLOBS = OUT / "blobs"
LIBDIR = REPO / "fast_ray_tracer_b200"

CFLAGS = ["-std=gnu11", "-O2", "-march=x86-64-v3", "-fPIC", "-w"]
WRAPS = ["-Wl,--wrap=render_multi", "-Wl,--wrap=intersect_world", "-Wl,--wrap=write_ppm_file", "-Wl,--wrap=write_png",
         "-Wl,--wrap=trace_photons"]

# name -> (yaml relative to the reference root, edit function on the parsed YAML list)


def _config(tree):
    for obj in tree:
        if isinstance(obj, dict) and obj.get("add") == "config":
            return obj
    return None


def _direct_only(tree):
    cfg = _config(tree)
    cfg["illumination"]["include-global"] = False
    cfg["illumination"]["global-illumination"]["photon-count"] = 0
    return tree


def _cache_size(tree, n):
    for obj in tree:
        if isinstance(obj, dict) and obj.get("add") == "light" and "corner" in obj:
            obj["cache-size"] = n
    return tree


def edit_cornell_gi(tree, photons=100000, caustics=False):
    """C5 at a size the CPU reference renders in seconds: the shipped GI configuration (global map + 8x8 final gather,
    kNN 200, r 0.1) with fewer photons and a 64-set light cache (small enough to keep as a fixture)."""
    cfg = _config(tree)
    cfg["illumination"]["global-illumination"]["photon-count"] = photons
    cfg["illumination"]["global-illumination"]["include-caustics"] = caustics
    return _cache_size(tree, 64)


ASSETS = OUT / "assets"
MAX_TEXTURE_DIM = 128


def edit_images(tree):
    """Image patterns: the generator shells out to ImageMagick for non-PNG files and the program opens them relative to
    its working directory.  Neither exists on the GPU box, so every image is converted (Pillow) or copied into
    oracle/_ref/assets/ and referenced by absolute path (/root/repo is valid on both boxes)."""
    from PIL import Image

    ASSETS.mkdir(parents=True, exist_ok=True)

    def walk(node):
        if isinstance(node, dict):
            if node.get("type") == "image" and "file" in node:
                src = REF / node["file"]
                dst = ASSETS / (Path(node["file"]).stem + ".png")
                if not dst.exists():
                    im = Image.open(src).convert("RGB")
                    if max(im.size) > MAX_TEXTURE_DIM:  # keep the scene blobs (texels as doubles) small enough to be fixtures
                        k = MAX_TEXTURE_DIM / max(im.size)
                        im = im.resize((max(1, round(im.size[0] * k)), max(1, round(im.size[1] * k))), Image.LANCZOS)
                    im.save(dst, format="PNG")
                node["file"] = "/root/repo/oracle/_ref/assets/" + dst.name
            for v in node.values():
                walk(v)
        elif isinstance(node, list):
            for v in node:
                walk(v)

    walk(tree)
    return tree


def edit_dof(tree):
    for obj in tree:
        if isinstance(obj, dict) and obj.get("add") == "camera":
            obj["aperture"] = {"jitter": True, "size": 0.06, "type": ["CIRCULAR_APERTURE", 1.0]}
            obj["usteps"] = 4
            obj["vsteps"] = 4
    return tree


def edit_none(tree):
    return tree


def edit_sibenik_surrogate(tree):
    """C4 stand-in: write the synthetic textured OBJ mesh + MTL + texture copies into oracle/_ref/assets/ first."""
    sys.path.insert(0, str(REPO / "oracle" / "scenes"))
    import make_sibenik_surrogate

    make_sibenik_surrogate.main(int(os.environ.get("FRT_SIBENIK_DETAIL", "3")))
    return tree


def edit_cornell_exact(tree):
    return _cache_size(_direct_only(tree), 1)


def edit_cornell_shipped(tree):
    return _direct_only(tree)


SCENES = {
    # C1
    "reflect_refract": ("scenes/reflect_refract/reflect_refract.yml", edit_none),
    # C2 exact (deterministic: one cached sample set) and as shipped (65535 sets picked with rand())
    "cornell_exact": ("scenes/cornell_box/cornell_box.yml", edit_cornell_exact),
    "cornell_shipped": ("scenes/cornell_box/cornell_box.yml", edit_cornell_shipped),
    # 64 cached sets: small enough to keep as a fixture, pins fast_ray_tracer_b200/lightcache.py bit for bit
    "cornell_cache64": ("scenes/cornell_box/cornell_box.yml", lambda t: _cache_size(_direct_only(t), 64)),
    # C5: photon-mapped (global map + final gather); and a variant that also fills and queries the caustic map
    "cornell_gi": ("scenes/cornell_box/cornell_box.yml", edit_cornell_gi),
    "cornell_gi_caustics": ("scenes/cornell_box/cornell_box.yml", lambda t: edit_cornell_gi(t, 100000, True)),
    # primitive / CSG / group coverage
    "group_test": ("scenes/group_test/group.yml", edit_none),
    "csg_test": ("scenes/test/test.yml", edit_none),
    "reflect_refract_test": ("scenes/reflect_refract_test/test.yml", edit_none),
    "checkered_torus": ("scenes/checkered_torus/checkered_torus.yml", edit_none),
    "checkered_sphere": ("scenes/checkered_sphere/checkered_sphere.yml", edit_none),
    "checkered_cube": ("scenes/checkered_cube/checkered_cube.yml", edit_none),
    "checkered_cylinder": ("scenes/checkered_cylinder/checkered_cylinder.yml", edit_none),
    "align_check_plane": ("scenes/align_check_plane/align_check_plane.yml", edit_none),
    "lens_test": ("scenes/lens_test/lens_test.yml", edit_none),
    "shadow_glamour_shot": ("scenes/shadow_glamour_shot/shadow_glamour_shot.yml", lambda t: _cache_size(t, 1)),
    # a scene of this repository: Perlin-perturbed / blended / nested / gradient patterns, cone, cylinder, circular area light
    "patterns_circle_light": (str(REPO / "oracle" / "scenes" / "patterns_circle_light.yml"), edit_none),
    # focal blur + jittered CMJ (stochastic in the reference: drand48): dof.yml with a circular aperture switched on
    "dof_blur": ("scenes/dof_test/dof.yml", lambda t: edit_dof(t)),
    # image textures (Ka / Kd / bump maps through planar and spherical uv maps)
    "bump_map_test": ("scenes/bump_map_test/bump_map_test.yml", edit_images),
    "texture_map_test": ("scenes/texture_map_test/texture_map_test.yml", edit_images),
    # C4 stand-in (sibenik.obj is not in the reference tree): textured + bump-mapped OBJ triangles under an area light
    "sibenik_surrogate": (str(REPO / "oracle" / "scenes" / "sibenik_surrogate.yml"), edit_sibenik_surrogate),
    # C3
    "teapot": ("scenes/teapot/teapot.yml", edit_none),
    "bounding_boxes": ("scenes/bounding_boxes/bounding_boxes.yml", edit_none),
}


# scenes that load an OBJ file (construct_group_from_obj_file)
OBJ_SCENES = {"teapot", "bounding_boxes", "sibenik_surrogate"}

# scenes whose blob would be too large to ship on every gpurun snapshot (the light cache alone is 157 MB)
NO_BLOB = {"cornell_shipped"}


def run(cmd, **kw):
    r = subprocess.run(cmd, **kw)
    if r.returncode != 0:
        raise SystemExit(f"command failed ({r.returncode}): {' '.join(map(str, cmd))}")
    return r


def reference_sources():
    srcs = sorted(p for p in (REF / "src").rglob("*.c") if "core_select" not in p.parts)
    return srcs


def obj_path(src: Path) -> Path:
    rel = src.relative_to(REF / "src")
    return OBJ / ("__".join(rel.with_suffix("").parts) + ".o")


def build_objects(force=False):
    OBJ.mkdir(parents=True, exist_ok=True)
    for src in reference_sources():
        o = obj_path(src)
        if force or not o.exists() or o.stat().st_mtime < src.stat().st_mtime:
            run(["gcc", *CFLAGS, "-I", str(REPO / "oracle" / "png_stub"), "-c", str(src), "-o", str(o)])
    hooks = OBJ / "ref_hooks.o"
    hsrc = REPO / "oracle" / "ref_hooks.c"
    if force or not hooks.exists() or hooks.stat().st_mtime < hsrc.stat().st_mtime:
        run(["gcc", *CFLAGS, "-I", str(REF), "-I", str(REPO / "oracle" / "png_stub"), "-c", str(hsrc), "-o", str(hooks)])
    objload = OBJ / "frt_objload.o"
    osrc = LIBDIR / "csrc" / "frt_objload.c"
    if force or not objload.exists() or objload.stat().st_mtime < osrc.stat().st_mtime:
        run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-Wall", "-I", str(REF), "-I", str(REPO / "oracle" / "png_stub"), "-c", str(osrc),
             "-o", str(objload)])
    shim = OBJ / "frt_shim.o"
    ssrc = LIBDIR / "csrc" / "frt_shim.c"
    abi = REPO / "include" / "frt_b200.h"  # the shim embeds FRT_ABI_VERSION and the struct layouts
    if force or not shim.exists() or shim.stat().st_mtime < max(ssrc.stat().st_mtime, abi.stat().st_mtime):
        run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-Wall", "-I", str(REF), "-I", str(REPO / "include"),
             "-I", str(REPO / "oracle" / "png_stub"), "-c", str(ssrc), "-o", str(shim)])


def build_objload_check():
    """oracle/_ref/objload_check: the reference's OBJ loader and the rebuilt one side by side (oracle/objload_check.c)."""
    exe = OUT / "objload_check"
    src = REPO / "oracle" / "objload_check.c"
    replaced = {"renderer__renderer.o", "renderer__photon_tracer.o"}
    host_objs = [str(obj_path(s)) for s in reference_sources() if obj_path(s).name not in replaced]
    deps = [src, OBJ / "frt_objload.o", OBJ / "frt_shim.o"]
    if not exe.exists() or exe.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        run(["gcc", *CFLAGS, "-I", str(REF), "-I", str(REPO / "oracle" / "png_stub"), "-o", str(exe), str(src), *host_objs,
             str(OBJ / "frt_shim.o"), str(OBJ / "frt_objload.o"), "-Wl,--wrap=construct_group_from_obj_file", "-L", str(LIBDIR), "-lfrt_b200",
             "-Wl,-rpath,$ORIGIN/../../fast_ray_tracer_b200", "-lm", "-lpthread", "-lz"])
    return exe


def build_pm_oracle():
    """oracle/_ref/libpm_ref.so: the reference's photon map (pm.c, unmodified) behind oracle/pm_oracle.c."""
    lib = OUT / "libpm_ref.so"
    src = REPO / "oracle" / "pm_oracle.c"
    pm = REF / "src" / "libs" / "photon_map" / "pm.c"
    if not lib.exists() or lib.stat().st_mtime < max(src.stat().st_mtime, pm.stat().st_mtime):
        run(["gcc", *CFLAGS, "-shared", "-I", str(REF), "-o", str(lib), str(src), str(pm), "-lm"])
    return lib


def build_canvas_oracle():
    """oracle/_ref/libcanvas_ref.so: the reference's canvas_pixel_at (canvas.c, unmodified, with its colour helpers)."""
    lib = OUT / "libcanvas_ref.so"
    src = REPO / "oracle" / "canvas_oracle.c"
    deps = [REF / "src" / "libs" / "canvas" / "canvas.c"] + sorted((REF / "src" / "color").glob("*.c")) + [REF / "src" / "libs" / "linalg" / "linalg.c"]
    deps = [p for p in deps if p.exists()]
    if not lib.exists() or lib.stat().st_mtime < max(p.stat().st_mtime for p in [src, *deps]):
        run(["gcc", *CFLAGS, "-shared", "-I", str(REF), "-I", str(REPO / "oracle" / "png_stub"), "-o", str(lib), str(src), *map(str, deps), "-lm", "-lz"])
    return lib


def generate_main(name: str) -> Path:
    import yaml

    rel, edit = SCENES[name]
    GEN.mkdir(parents=True, exist_ok=True)
    with open(REF / rel) as f:
        tree = yaml.safe_load(f)
    tree = edit(copy.deepcopy(tree))
    yml = GEN / f"{name}.yml"
    with open(yml, "w") as f:
        yaml.safe_dump(tree, f, default_flow_style=None, sort_keys=False)
    main_c = GEN / f"{name}.c"
    # the generator does flat imports and the scenes use reference-root-relative asset paths: run it from there
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    with open(main_c, "w") as f:
        run([sys.executable, str(REF / "yaml_parser" / "yaml_parser.py"), str(yml)], cwd=str(REF), stdout=f, env=env)
    return main_c


def build_scene(name: str, force=False):
    main_c = generate_main(name)
    main_o = OBJ / f"main__{name}.o"
    run(["gcc", *CFLAGS, "-I", str(REF), "-I", str(REPO / "oracle" / "png_stub"), "-c", str(main_c), "-o", str(main_o)])
    all_objs = [str(obj_path(s)) for s in reference_sources()]
    hooks = str(OBJ / "ref_hooks.o")
    ref_bin = OUT / f"{name}_ref"
    run(["gcc", "-o", str(ref_bin), str(main_o), *all_objs, hooks, *WRAPS, "-lm", "-lpthread", "-lz"])
    replaced = {"renderer__renderer.o", "renderer__photon_tracer.o"}
    host_objs = [o for o in all_objs if Path(o).name not in replaced]
    b200_bin = OUT / f"{name}_b200"
    run(["gcc", "-o", str(b200_bin), str(main_o), *host_objs, str(OBJ / "frt_shim.o"), hooks, *WRAPS,
         "-L", str(LIBDIR), "-lfrt_b200", "-Wl,-rpath,$ORIGIN/../../fast_ray_tracer_b200", "-lm", "-lpthread", "-lz"])
    if name in OBJ_SCENES:
        # the drop-in with the rebuilt OBJ parse as well (csrc/frt_objload.c behind --wrap=construct_group_from_obj_file;
        # FRT_OBJLOAD=ref sends the call back to the reference's loader): tests/test_objload.py compares the two blobs
        run(["gcc", "-o", str(OUT / f"{name}_b200obj"), str(main_o), *host_objs, str(OBJ / "frt_shim.o"), str(OBJ / "frt_objload.o"), hooks,
             *WRAPS, "-Wl,--wrap=construct_group_from_obj_file", "-L", str(LIBDIR), "-lfrt_b200",
             "-Wl,-rpath,$ORIGIN/../../fast_ray_tracer_b200", "-lm", "-lpthread", "-lz"])
    return ref_bin, b200_bin


def dump_blob(name: str, extra_env=None, suffix="") -> Path:
    """Run the drop-in build with FRT_DUMP_ONLY=1 from the reference root (asset paths) and keep the blob."""
    BLOBS.mkdir(parents=True, exist_ok=True)
    blob = BLOBS / f"{name}{suffix}.frt"
    env = dict(os.environ, FRT_DUMP_SCENE=str(blob), FRT_DUMP_ONLY="1", FRT_SKIP_PPM="1")
    env.update(extra_env or {})
    run([str(OUT / f"{name}_b200")], cwd=str(REF), env=env, stdout=subprocess.DEVNULL)
    return blob


def run_reference(name: str, canvas_out: Path, extra_env=None, threads=None):
    env = dict(os.environ, FRT_CANVAS_OUT=str(canvas_out), FRT_SKIP_PPM="1")
    env["FRT_REF_THREADS"] = str(threads or os.cpu_count() or 1)
    env.update(extra_env or {})
    r = run([str(OUT / f"{name}_ref")], cwd=str(REF), env=env, stdout=subprocess.PIPE, text=True)
    info = {}
    for line in r.stdout.splitlines():
        if line.startswith("FRT_"):
            k, _, v = line.partition(" ")
            info[k] = v
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scenes", nargs="*", default=[], help="scene names (default: all)")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--no-blobs", action="store_true")
    args = ap.parse_args()
    if not REF.exists():
        print(f"reference tree {REF} not present: keeping the prebuilt oracle/_ref as is")
        return 0
    if not (LIBDIR / "libfrt_b200.so").exists():
        raise SystemExit("build fast_ray_tracer_b200/libfrt_b200.so first (python -m fast_ray_tracer_b200.build)")
    names = args.scenes or list(SCENES)
    build_objects(force=args.force)
    build_pm_oracle()
    build_canvas_oracle()
    build_objload_check()
    for name in names:
        build_scene(name, force=args.force)
        if not args.no_blobs and name not in NO_BLOB:
            dump_blob(name)
        print(f"built {name}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
