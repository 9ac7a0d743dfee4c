"""Host logic: the bounding groups the upload inserts over long runs of triangle children (frt_leafruns.h) leave the
reference's leaf order -- which its shadow rays depend on, group.c:105-123 -- and every original node untouched."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN, REPO

GROUP, CSG, TRI, STRI = 9, 8, 7, 4
RUN_MIN, RUN_LEAF = 6, 4
NODE_DT = np.dtype([("type", "i4"), ("skip", "i4"), ("parent", "i4"), ("xform", "i4"), ("material", "i4"), ("param", "i4"),
                    ("csg_op", "i4"), ("right", "i4"), ("bmin", "f8", 3), ("bmax", "f8", 3)])

BLOBS = [GOLDEN / "teapot.frt", GOLDEN / "group_test.frt", GOLDEN / "cornell_exact_200.frt",
         REPO / "oracle" / "_ref" / "blobs" / "bounding_boxes.frt", REPO / "oracle" / "_ref" / "blobs" / "sibenik_surrogate.frt"]


def as_array(nodes, n):
    from fast_ray_tracer_b200.api import frt_node

    assert C.sizeof(frt_node) == NODE_DT.itemsize
    raw = np.ctypeslib.as_array(C.cast(nodes, C.POINTER(C.c_byte)), shape=(n * NODE_DT.itemsize,))
    return raw.view(NODE_DT).copy()


@pytest.mark.parametrize("path", BLOBS, ids=lambda p: p.stem)
def test_inserted_groups_keep_the_leaf_order_and_bound_their_triangles(frt, path):
    if not path.exists():
        pytest.skip(f"{path.name} is built by oracle/build_ref.py")
    desc = frt.SceneDesc.load(path)
    check_tree(desc)


def check_tree(desc):
    """Every invariant of the tree the upload builds; returns (number of inserted groups, the new node array)."""
    d = desc.c
    old = as_array(d.nodes, d.n_nodes)
    nodes, roots = desc.tree_with_runs()
    new = as_array(nodes, len(nodes))
    n = len(new)
    prm = np.ctypeslib.as_array(d.prim_params, shape=(d.n_prim_params,)) if d.n_prim_params > 0 else np.zeros(1)

    # the tree is a well-formed pre-order list
    idx = np.arange(n)
    assert np.all(new["skip"] > idx) and np.all(new["skip"] <= n)
    assert np.all(new["parent"] < idx)
    has_parent = new["parent"] >= 0
    assert np.all(new["skip"][has_parent] <= new["skip"][new["parent"][has_parent]])
    leaf = new["type"] < CSG
    assert np.all(new["skip"][leaf] == idx[leaf] + 1)
    assert roots[0] == 0 and len(roots) == d.n_roots

    # leaves: same sequence, same contents
    old_leaf = old[old["type"] < CSG]
    for f in ("type", "xform", "material", "param"):
        assert np.array_equal(new[leaf][f], old_leaf[f]), f

    # original inner nodes: same sequence, same boxes; what is left over was inserted
    inner = np.where(~leaf)[0]
    old_inner = old[old["type"] >= CSG]
    key_old = np.concatenate([old_inner["bmin"], old_inner["bmax"]], axis=1)
    key_new = np.concatenate([new[inner]["bmin"], new[inner]["bmax"]], axis=1)
    j = 0
    inserted = []  # an inserted box is padded, so it never equals the box the reference computed for the same triangles
    for k, i in enumerate(inner):
        if j < len(old_inner) and new["type"][i] == old_inner["type"][j] and np.array_equal(key_new[k], key_old[j], equal_nan=True):
            j += 1
        else:
            inserted.append(i)
    assert j == len(old_inner)
    assert n == d.n_nodes + len(inserted)

    # an inserted group holds nothing but triangles of one transform, and its box holds their vertices
    for i in inserted:
        sub = new[i + 1:new["skip"][i]]
        tl = sub[sub["type"] < CSG]
        assert len(tl) >= 1 and np.all((tl["type"] == TRI) | (tl["type"] == STRI))
        assert np.all(sub[sub["type"] >= CSG]["type"] == GROUP)
        assert np.all(tl["xform"] == new["xform"][i])
        v = np.stack([prm[p:p + 9].reshape(3, 3) for p in tl["param"]])
        assert np.all(v.min(axis=(0, 1)) >= new["bmin"][i]) and np.all(v.max(axis=(0, 1)) <= new["bmax"][i])

    # no long run of triangle children is left anywhere outside a CSG
    under_csg = np.zeros(n, dtype=bool)
    for i in range(n):
        p = new["parent"][i]
        under_csg[i] = p >= 0 and (under_csg[p] or new["type"][p] == CSG)
    longest = 0
    for g in np.where((new["type"] == GROUP) & ~under_csg)[0]:
        run, c = 0, g + 1
        prev_xf = None
        while c < new["skip"][g]:
            if new["type"][c] in (TRI, STRI) and (prev_xf is None or prev_xf == new["xform"][c]):
                run += 1
            elif new["type"][c] in (TRI, STRI):
                run = 1
            else:
                run = 0
            prev_xf = new["xform"][c] if new["type"][c] in (TRI, STRI) else None
            longest = max(longest, run)
            c = new["skip"][c]
    assert longest < RUN_MIN
    return len(inserted), new


def test_runs_under_a_csg_or_of_mixed_transforms_are_left_alone(frt):
    """A hand-made tree: a group that is one run of 8 triangles (split without a group of its own), a CSG whose operands hold
    triangles (untouched: the CSG's crossing lists are per operand), triangles of alternating transforms (no run), and a run of
    7 next to a sphere (gets a group of its own)."""
    from fast_ray_tracer_b200.api import frt_node, frt_xform

    SPHERE = 5
    desc = frt.SceneDesc.load(GOLDEN / "teapot.frt")  # materials, camera and the rest of a valid description
    d = desc.c
    rng = np.random.default_rng(3)
    tris = []

    def tri(xform=0):
        p = rng.uniform(-1, 1, (3, 3))
        rec = np.zeros(34)
        rec[0:9] = p.reshape(-1)
        rec[9:12], rec[12:15] = p[1] - p[0], p[2] - p[0]
        tris.append(rec)
        return dict(type=TRI, xform=xform, material=0, param=34 * (len(tris) - 1))

    def group(children, xform=0):
        return dict(type=GROUP, xform=xform, material=-1, param=-1, children=children)

    tree = group([
        group([tri() for _ in range(8)]),
        dict(type=CSG, xform=0, material=-1, param=-1, children=[group([tri() for _ in range(7)]), tri()]),
        group([tri(k % 2) for k in range(8)]),
        *[tri() for _ in range(7)],
        dict(type=SPHERE, xform=0, material=0, param=-1),
    ])
    flat = []

    def emit(node, parent):
        i = len(flat)
        flat.append(None)
        kids = [emit(c, i) for c in node.get("children", [])]
        flat[i] = dict(node, parent=parent, skip=len(flat), right=kids[1] if node["type"] == CSG else -1)
        return i

    emit(tree, -1)
    arr = (frt_node * len(flat))()
    for i, nd in enumerate(flat):
        arr[i].type, arr[i].skip, arr[i].parent, arr[i].xform = nd["type"], nd["skip"], nd["parent"], nd["xform"]
        arr[i].material, arr[i].param, arr[i].csg_op, arr[i].right = nd["material"], nd["param"], 0, nd["right"]
        for k in range(3):
            arr[i].bbox_min[k], arr[i].bbox_max[k] = -2.0, 2.0
    prm = np.concatenate(tris)
    xf = (frt_xform * 2)()
    for j in range(2):
        for k, v in enumerate([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]):
            xf[j].inv[k] = float(v)
    xf[1].inv[3] = 0.5  # a translation: not the identity
    roots = (C.c_int32 * 1)(0)
    d.nodes, d.n_nodes, d.roots, d.n_roots = arr, len(flat), roots, 1
    d.xforms, d.n_xforms = xf, 2
    d.prim_params, d.n_prim_params = prm.ctypes.data_as(C.POINTER(C.c_double)), len(prm)
    desc._owned = False  # the arrays above are ours: the loader's free must not see them
    inserted, new = check_tree(desc)
    # the 8-run that fills its group is split without a group of its own (>= 2 groups), the 7-run next to the sphere gets one
    # around it and >= 2 below; how the ordered surface-area splits fall depends on the triangles
    assert 5 <= inserted <= 12
    csg = int(np.where(new["type"] == CSG)[0][0])
    sub = new[csg:new["skip"][csg]]
    assert len(sub) == 1 + 1 + 7 + 1  # the CSG, its left group, seven triangles, the right triangle: nothing inserted
    mixed = csg + len(sub)  # the group of alternating transforms follows the CSG: untouched as well
    assert new["type"][mixed] == GROUP and new["skip"][mixed] - mixed == 1 + 8
    assert new["type"][-1] == SPHERE and new["parent"][-1] == 0


def test_small_scenes_are_left_alone(frt):
    desc = frt.SceneDesc.load(GOLDEN / "cornell_exact_200.frt")
    nodes, _ = desc.tree_with_runs()
    assert len(nodes) == desc.c.n_nodes
