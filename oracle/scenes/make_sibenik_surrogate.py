#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- stand-in for BASELINE.json configs[3] (scenes/sibenik).

The reference ships scenes/sibenik/sibenik.yml, its MTL file and its five texture / bump PNGs, but NOT sibenik.obj
(SURVEY.md 8d).  This script writes a synthetic cathedral-like OBJ mesh with the same ingredients -- `vt` texture
coordinates on every face, smooth normals on the columns, materials bound through an MTL file to the reference's own
sibenik PNGs (map_Ka / map_Kd / map_bump through the TRIANGLE_UV_MAP path of obj_loader.c:60-98) -- into
oracle/_ref/assets/, which travels to the GPU box.  It is a SURROGATE: same code paths, not the same picture.

    nave      floor (marble, bump), two side walls and an apse wall (stone, bump), a barrel vault (stone)
    columns   two rows of tessellated, smooth-shaded columns (column texture)
    windows   glass quads in the side walls (Tf 0.5, no shadow test needed: they are ordinary casters in the reference)
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference")
ASSETS = REPO / "oracle" / "_ref" / "assets"
TEX = 256  # texture edge on the GPU box (the originals are 512): keeps the flattened scene (texels as doubles) small


class Mesh:
    def __init__(self):
        self.v, self.vt, self.vn, self.lines = [], [], [], []

    def vert(self, p):
        self.v.append(p)
        return len(self.v)

    def tex(self, t):
        self.vt.append(t)
        return len(self.vt)

    def nrm(self, n):
        self.vn.append(n)
        return len(self.vn)

    def grid(self, name, mtl, origin, du, dv, nu, nv, tile_u, tile_v, fn=None):
        """A (nu x nv)-cell patch: point(i, j) = fn(origin + i du + j dv) (fn bends it), uv tiled tile_u x tile_v times."""
        self.lines.append(f"g {name}")
        self.lines.append(f"usemtl {mtl}")
        idx = {}
        for j in range(nv + 1):
            for i in range(nu + 1):
                p = [origin[k] + du[k] * i / nu + dv[k] * j / nv for k in range(3)]
                if fn:
                    p = fn(p, i / nu, j / nv)
                idx[i, j] = (self.vert(p), self.tex((tile_u * i / nu, tile_v * j / nv)))
        for j in range(nv):
            for i in range(nu):
                a, b, c, d = idx[i, j], idx[i + 1, j], idx[i + 1, j + 1], idx[i, j + 1]
                self.lines.append(f"f {a[0]}/{a[1]} {b[0]}/{b[1]} {c[0]}/{c[1]} {d[0]}/{d[1]}")  # fan-triangulated by the loader

    def column(self, name, mtl, cx, cz, y0, y1, r, seg, rings):
        self.lines.append(f"g {name}")
        self.lines.append(f"usemtl {mtl}")
        idx = {}
        for j in range(rings + 1):
            y = y0 + (y1 - y0) * j / rings
            rr = r * (1.0 + 0.25 * (abs(2 * j / rings - 1) ** 6))  # flared base and capital
            for i in range(seg + 1):
                a = 2 * math.pi * i / seg
                idx[i, j] = (self.vert((cx + rr * math.cos(a), y, cz + rr * math.sin(a))), self.tex((2.0 * i / seg, 3.0 * j / rings)),
                             self.nrm((math.cos(a), 0.0, math.sin(a))))
        for j in range(rings):
            for i in range(seg):
                a, b, c, d = idx[i, j], idx[i, j + 1], idx[i + 1, j + 1], idx[i + 1, j]
                self.lines.append("f " + " ".join(f"{q[0]}/{q[1]}/{q[2]}" for q in (a, b, c, d)))

    def write(self, path: Path, mtl_path: Path):
        with open(path, "w") as f:
            f.write("# synthetic stand-in for scenes/sibenik/sibenik.obj (oracle/scenes/make_sibenik_surrogate.py)\n")
            f.write(f"mtllib {mtl_path}\n")
            for p in self.v:
                f.write("v %.9g %.9g %.9g\n" % tuple(p))
            for t in self.vt:
                f.write("vt %.9g %.9g 0\n" % tuple(t))
            for n in self.vn:
                f.write("vn %.9g %.9g %.9g\n" % tuple(n))
            f.write("\n".join(self.lines) + "\n")


def build(detail: int = 1) -> Mesh:
    m = Mesh()
    W, H, L = 4.0, 5.0, 14.0  # half width, wall height, length (z from -L/2 to L/2)
    n = 12 * detail
    m.grid("floor", "pod", (-W, 0, -L / 2), (2 * W, 0, 0), (0, 0, L), 2 * n, 3 * n, 6, 10)
    m.grid("wall_left", "kamen_zid", (-W, 0, L / 2), (0, 0, -L), (0, H, 0), 3 * n, n, 6, 2)
    m.grid("wall_right", "kamen_zid", (W, 0, -L / 2), (0, 0, L), (0, H, 0), 3 * n, n, 6, 2)
    m.grid("apse", "kamen_zid", (-W, 0, L / 2), (2 * W, 0, 0), (0, H + W, 0), 2 * n, 2 * n, 4, 4)
    m.grid("entry", "kamen_zid", (W, 0, -L / 2), (-2 * W, 0, 0), (0, H + W, 0), 2 * n, 2 * n, 4, 4)

    def vault(p, s, t):  # bend the flat strip x in [-W, W] into a half cylinder above the walls
        a = math.pi * s
        return [-W * math.cos(a), H + W * math.sin(a) * 0.75, p[2]]

    m.grid("vault", "kamen_zid", (-W, H, -L / 2), (2 * W, 0, 0), (0, 0, L), 2 * n, 3 * n, 4, 8, fn=vault)
    for side in (-1, 1):
        for k in range(5):
            z = -L / 2 + (k + 0.5) * L / 5
            m.column(f"column_{'l' if side < 0 else 'r'}{k}", "stupovi", side * (W - 1.3), z, 0.0, H - 0.4, 0.32, 12 * detail, 8 * detail)
            # a window in the wall behind every column
            x = side * (W - 0.02)
            m.grid(f"window_{'l' if side < 0 else 'r'}{k}", "staklo", (x, 2.0, z - 0.5 * side), (0, 0, 1.0 * side), (0, 2.0, 0), 2, 4, 1, 1)
    return m


MTL = """# materials of scenes/sibenik/sibenik.mtl (the ones the surrogate mesh uses), texture paths re-pointed at copies of the
# reference's own PNGs under oracle/_ref/assets/ (absolute: valid in this container and on the GPU box)
newmtl pod
	Ns 8.0
	Ni 1.0
	d 1.0
	Tf 1.0 1.0 1.0
	illum 2
	Ka 0.05 0.05 0.05
	Kd 0.70 0.70 0.70
	Ks 0.15 0.15 0.15
	map_Ka {a}/sibenik_mramor6x6.png
	map_Kd {a}/sibenik_mramor6x6.png
	map_bump {a}/sibenik_mramor6x6-bump.png

newmtl kamen_zid
	Ns 8.0
	Ni 1.0
	d 1.0
	Tf 1.0 1.0 1.0
	illum 2
	Ka 0.05 0.05 0.05
	Kd 0.734118 0.730588 0.674118
	Ks 0.0 0.0 0.0
	map_Ka {a}/sibenik_kamen.png
	map_Kd {a}/sibenik_kamen.png
	map_bump {a}/sibenik_kamen-bump.png

newmtl stupovi
	Ns 8.0
	Ni 1.0
	d 1.0
	Tf 1.0 1.0 1.0
	illum 2
	Ka 0.05 0.05 0.05
	Kd 0.734118 0.730588 0.674118
	Ks 0.0 0.0 0.0
	map_Ka {a}/sibenik_KAMEN-stup.png
	map_Kd {a}/sibenik_KAMEN-stup.png

newmtl staklo
	Ns 256.0
	Ni 1.0
	Tf 0.5 0.5 0.5
	illum 6
	Ka 0.0 0.0 0.0
	Kd 0.0 0.0 0.0
	Ks 0.1 0.1 0.1
"""


def main(detail: int = 1):
    from PIL import Image

    ASSETS.mkdir(parents=True, exist_ok=True)
    for name in ("mramor6x6", "mramor6x6-bump", "kamen", "kamen-bump", "KAMEN-stup"):
        dst = ASSETS / f"sibenik_{name}.png"
        if not dst.exists():
            im = Image.open(REF / "scenes" / "sibenik" / f"{name}.png").convert("RGB")
            k = TEX / max(im.size)
            im.resize((max(1, round(im.size[0] * k)), max(1, round(im.size[1] * k))), Image.LANCZOS).save(dst, format="PNG")
    mtl = ASSETS / "sibenik_surrogate.mtl"
    mtl.write_text(MTL.format(a="/root/repo/oracle/_ref/assets"))
    mesh = build(detail)
    obj = ASSETS / "sibenik_surrogate.obj"
    mesh.write(obj, Path("/root/repo/oracle/_ref/assets") / mtl.name)
    tris = sum(len(l.split()) - 3 for l in mesh.lines if l.startswith("f "))
    print(f"{obj}: {len(mesh.v)} vertices, {tris} triangles")
    return obj


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
