"""Host logic: scene blobs written by the C shim (flattened reference World) load back intact."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, golden_names

GROUP, CSG, CUBE, SPHERE = 9, 8, 1, 5


def nodes_of(desc):
    d = desc.c
    return [d.nodes[i] for i in range(d.n_nodes)]


@pytest.mark.parametrize("name", golden_names())
def test_blob_is_well_formed(frt, name):
    desc = frt.SceneDesc.load(GOLDEN / f"{name}.frt")
    d = desc.c
    assert d.n_nodes > 0 and d.n_roots >= 1 and d.n_xforms >= 1
    ident = np.array(d.xforms[0].inv[:]).reshape(3, 4)
    assert np.array_equal(ident, np.eye(4)[:3])
    for i, n in enumerate(nodes_of(desc)):
        assert i < n.skip <= d.n_nodes
        assert 0 <= n.xform < d.n_xforms
        assert -1 <= n.parent < i
        if n.type < CSG:
            assert 0 <= n.material < d.n_materials
        if n.type == CSG:
            assert i + 1 < n.right < n.skip
    cam = desc.camera
    assert cam.hsize > 0 and cam.vsize > 0 and d.n_pixel_samples == 2 * cam.usteps * cam.vsteps


def test_cornell_tree_matches_the_reference_divide(frt):
    """SURVEY.md 3.5 fixture: the Cornell tree after divide(1), child order exactly as the reference holds it."""
    desc = frt.SceneDesc.load(GOLDEN / "cornell_exact_200.frt")
    ns = nodes_of(desc)

    def children(i):
        out, j = [], i + 1
        while j < ns[i].skip:
            out.append(j)
            j = ns[j].skip
        return out

    root = desc.c.roots[0]
    top = children(root)
    assert [ns[i].type for i in top] == [GROUP, GROUP, CUBE, CUBE, CUBE, CUBE]
    left, right = top[0], top[1]
    assert [ns[i].type for i in children(left)] == [GROUP, CUBE]
    assert [ns[i].type for i in children(children(left)[0])] == [SPHERE]
    assert [ns[i].type for i in children(right)] == [CSG, CUBE]
    csg = children(right)[0]
    assert ns[csg].csg_op == 2  # difference
    inner = children(csg)
    assert [ns[i].type for i in inner] == [CSG, CUBE] and ns[inner[0]].csg_op == 0  # union
    assert [ns[i].type for i in children(inner[0])] == [CUBE, CUBE]
    # divide-created groups are identity; every primitive carries its own transform
    for i, n in enumerate(ns):
        if n.type == GROUP:
            assert n.xform == 0
        elif n.type < CSG:
            assert n.xform != 0
    light = desc.c.lights[0]
    assert light.num_samples == 100 and light.cache_len == 1 and desc.c.n_light_points == 100


def test_save_load_round_trip(frt, tmp_path):
    a = frt.SceneDesc.load(GOLDEN / "teapot.frt")
    a.save(tmp_path / "copy.frt")
    assert (tmp_path / "copy.frt").read_bytes() == (GOLDEN / "teapot.frt").read_bytes()
    b = frt.SceneDesc.load(tmp_path / "copy.frt")
    assert b.c.n_nodes == a.c.n_nodes and b.c.n_prim_params == a.c.n_prim_params
    n = a.c.n_prim_params
    assert np.array_equal(np.ctypeslib.as_array(a.c.prim_params, (n,)), np.ctypeslib.as_array(b.c.prim_params, (n,)))


def test_load_rejects_garbage(frt, tmp_path):
    p = tmp_path / "bad.frt"
    p.write_bytes(b"\0" * 4096)
    with pytest.raises(frt.FrtError):
        frt.SceneDesc.load(p)
    with pytest.raises(frt.FrtError):
        frt.SceneDesc.load(tmp_path / "missing.frt")


def test_set_resolution_follows_the_reference_camera(frt):
    desc = frt.SceneDesc.load(GOLDEN / "reflect_refract.frt")
    cam = desc.camera
    hw, ps = cam.half_width, cam.pixel_size
    desc.set_resolution(800, 400)  # same aspect: half extents unchanged, pixel halves
    assert cam.half_width == hw and abs(cam.pixel_size - ps / 2) < 1e-15
    desc.set_resolution(200, 400)  # portrait: camera.c:127-130
    assert abs(cam.half_height - hw) < 1e-15 and abs(cam.half_width - hw * 0.5) < 1e-15


def test_malformed_descriptions_are_rejected_before_any_device_work(frt):
    """validate_desc bounds-checks every index the kernels dereference (FRT_ERR_ARG, also on a host without a GPU)."""
    import pytest

    def load():
        return frt.SceneDesc.load(GOLDEN / "texture_map_test.frt")

    d = load()
    d.c.materials[0].map_Kd = d.c.n_patterns
    with pytest.raises(frt.FrtError, match="material 0"):
        frt.Scene(d)
    d = load()
    d.c.textures[0].texel_offset = d.c.n_texels
    with pytest.raises(frt.FrtError, match="texture 0"):
        frt.Scene(d)
    d = load()
    for i in range(d.c.n_patterns):
        if d.c.patterns[i].type == 9:  # uv texture
            d.c.patterns[i].i[0] = d.c.n_textures + 3
            break
    with pytest.raises(frt.FrtError, match="pattern"):
        frt.Scene(d)
    d = load()
    d.c.n_materials = -1
    with pytest.raises(frt.FrtError, match="negative"):
        frt.Scene(d)
    d = frt.SceneDesc.load(GOLDEN / "teapot.frt")
    for i in range(d.c.n_nodes):
        if d.c.nodes[i].type in (4, 7):  # a triangle whose parameters would run past the pool
            d.c.nodes[i].param = d.c.n_prim_params - 5
            break
    with pytest.raises(frt.FrtError, match="parameter offset"):
        frt.Scene(d)


def test_blob_with_a_negative_count_is_refused(frt, tmp_path):
    import struct

    import pytest

    raw = bytearray((GOLDEN / "csg_test.frt").read_bytes())
    # header: uint64 magic, int32 version, then the int32 counts (n_nodes first), then four int64 counts
    for offset, fmt, value in ((12, "<i", -5), (12 + 7 * 4 + 4 + 16, "<q", -(1 << 40)), (12 + 7 * 4 + 4 + 8, "<q", 1 << 60)):
        bad_raw = bytearray(raw)
        struct.pack_into(fmt, bad_raw, offset, value)
        bad = tmp_path / "bad.frt"
        bad.write_bytes(bytes(bad_raw))
        with pytest.raises(frt.FrtError, match="negative or impossible|truncated"):
            frt.SceneDesc.load(bad)
