"""Host logic: the rebuilt OBJ parse (fast_ray_tracer_b200/csrc/frt_objload.c) builds the shape tree
the reference's construct_group_from_obj_file (src/libs/obj_loader/obj_loader.c:446) builds.

oracle/_ref/objload_check (oracle/objload_check.c, built by oracle/build_ref.py from the reference's own objects) loads a
file with both and compares the trees field by field; oracle/_ref/<scene>_b200obj is the drop-in program with the rebuilt
parse linked in, whose flattened scene must equal the one the reference's loader leads to, byte for byte."""
import os
import random
import re
import subprocess
from pathlib import Path

import pytest

from conftest import REPO

REF = Path(os.environ.get("FRT_REFERENCE_ROOT", "/root/reference"))
CHECK = REPO / "oracle" / "_ref" / "objload_check"
LINE = re.compile(rb"OBJLOAD groups (\d+) triangles (\d+) differences (\d+) reference_ms ([0-9.]+) rebuilt_ms ([0-9.]+)")


def check(path, cwd=None):
    if not CHECK.exists():
        pytest.skip("oracle/_ref/objload_check is built by oracle/build_ref.py")
    r = subprocess.run([str(CHECK), str(path)], cwd=cwd, capture_output=True, timeout=300)
    m = LINE.search(r.stdout)
    assert m is not None, r.stdout[-2000:] + r.stderr[-2000:]
    groups, tris, diff = int(m.group(1)), int(m.group(2)), int(m.group(3))
    assert diff == 0 and r.returncode == 0, r.stdout[-2000:]
    return groups, tris, float(m.group(4)), float(m.group(5))


def number(rng):
    """Decimal spellings sscanf's %lf accepts: plain, signed, exponents, many digits, no integer / no fraction part."""
    v = rng.uniform(-50, 50)
    return rng.choice([
        f"{v:.6f}", f"{v:.3f}", f"{v:.17g}", f"{v:+.4f}", f"{v:.5e}", f"{v:.12E}", f"{int(v)}", f"{int(v)}.", f".{rng.randrange(10 ** 6):06d}",
        f"{v:.20f}", f"000{abs(v):.2f}", f"{v * 1e-30:.8e}", f"{v * 1e25:.3f}", f"{rng.randrange(10 ** 17)}e-12",
    ])


def synthetic_obj(rng, n_vertices=400, n_faces=600, crlf=False):
    eol = "\r\n" if crlf else "\n"
    out = ["# synthetic", "o thing"]
    for _ in range(n_vertices):
        out.append("v " + " ".join(number(rng) for _ in range(3)))
    for _ in range(n_vertices):
        out.append("vn " + "  ".join(number(rng) for _ in range(3)))
    for _ in range(n_vertices):
        out.append("vt " + " ".join(number(rng) for _ in range(rng.choice([2, 3]))))  # two components: the third stays 0
    names = ["a", "b", "a", "c", "##default_group", "b"]
    for f in range(n_faces):
        if f % 97 == 0:
            out.append(f"g {names[(f // 97) % len(names)]}")
        k = rng.choice([3, 3, 3, 4, 5, 9])
        idx = [rng.randrange(1, n_vertices + 1) for _ in range(k)]
        style = rng.choice(["v", "v/t", "v/t/n", "v//n"])
        toks = []
        for i in idx:
            t, n = rng.randrange(1, n_vertices + 1), rng.randrange(1, n_vertices + 1)
            toks.append({"v": f"{i}", "v/t": f"{i}/{t}", "v/t/n": f"{i}/{t}/{n}", "v//n": f"{i}//{n}"}[style])
        sep = rng.choice([" ", "  ", "\t", " \t "])
        # a blank in front of "\r\n" makes the reference parse the "\r\n" token as a vertex (index 0 - 1: it reads in front of its
        # array and crashes or not); trailing blanks are only generated for "\n" files
        out.append("f " + sep.join(toks) + ("" if crlf else rng.choice(["", " ", "\t"])))
    out += ["s off", "usemtl nothing_by_that_name", "", "f 1 2 3", "vp 0.1 0.2", "v 1 2 3"]  # unknown lines, a vertex after the faces
    long_face = "f " + " ".join(str(rng.randrange(1, n_vertices + 1)) for _ in range(400))  # > 1023 characters: split like fgets does
    out.append(long_face)
    return eol.join(out) + eol


@pytest.mark.parametrize("seed,crlf", [(1, False), (2, False), (3, True)])
def test_synthetic_files_give_the_reference_tree(tmp_path, seed, crlf):
    rng = random.Random(seed)
    path = tmp_path / f"synthetic_{seed}.obj"
    path.write_text(synthetic_obj(rng, crlf=crlf), newline="")
    groups, tris, _, _ = check(path)
    assert groups >= 4 and tris > 600


@pytest.mark.parametrize("rel", ["scenes/teapot/teapot_low.obj", "scenes/bounding_boxes/dragon.obj"])
def test_reference_meshes_give_the_reference_tree(rel):
    if not (REF / rel).exists():
        pytest.skip("the reference tree is not present")
    groups, tris, ref_ms, own_ms = check(REF / rel, cwd=REF)
    assert tris >= 240
    print(f"{rel}: {tris} triangles, reference {ref_ms:.1f} ms, rebuilt {own_ms:.1f} ms")


def test_textured_mesh_with_materials_gives_the_reference_tree():
    obj = REPO / "oracle" / "_ref" / "assets" / "sibenik_surrogate.obj"
    if not obj.exists():
        pytest.skip("oracle/_ref/assets is written by oracle/build_ref.py")
    groups, tris, _, _ = check(obj)  # mtllib, usemtl, map_Kd / map_bump, vt, 26 named groups
    assert groups > 20 and tris > 80000


@pytest.mark.parametrize("scene", ["teapot", "bounding_boxes"])
def test_drop_in_program_flattens_to_the_same_scene(tmp_path, scene):
    exe = REPO / "oracle" / "_ref" / f"{scene}_b200obj"
    if not exe.exists() or not REF.exists():
        pytest.skip("needs oracle/_ref/<scene>_b200obj and the reference tree (asset paths)")
    blobs = {}
    for mode in ("ref", "own"):
        blob = tmp_path / f"{scene}_{mode}.frt"
        env = dict(os.environ, FRT_OBJLOAD=mode, FRT_DUMP_SCENE=str(blob), FRT_DUMP_ONLY="1", FRT_SKIP_PPM="1", FRT_WARM="0")
        subprocess.run([str(exe)], cwd=str(REF), env=env, stdout=subprocess.DEVNULL, check=True, timeout=600)
        blobs[mode] = blob.read_bytes()
    assert blobs["ref"] == blobs["own"]
